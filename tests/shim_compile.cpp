// Compile-and-link check of the header-only C++ shim (include/lgs/registration.hpp) against liblgs_b200.so.
// Mirrors the reference's call sequence (LSM:149,162-172; PPF:118-120,136-139).  Runs the calls only when a GPU exists.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "lgs/registration.hpp"

int main(int argc, char** argv) {
  const bool run = argc > 1 && std::strcmp(argv[1], "--run") == 0;
  if (!run) {
    std::printf("shim compiled; version %s\n", lgs_version());
    return 0;
  }
  auto cloud = std::make_shared<lgs::PointCloud>();
  for (int i = 0; i < 4000; i++) {
    float a = 0.01f * i;
    cloud->push_back(lgs::PointXYZI{10.0f * std::cos(a), 10.0f * std::sin(a), 0.001f * (i % 97), 1.0f, float(i % 255), 0, 0, 0});
  }
  lgs::VoxelGrid vg;
  vg.setLeafSize(0.2f, 0.2f, 0.2f);
  vg.setInputCloud(cloud);
  lgs::PointCloud filtered;
  vg.filter(filtered);
  lgs::StatisticalOutlierRemoval sor;  // PPF:132-140, on the voxel grid's output
  sor.setMeanK(30);
  sor.setStddevMulThresh(1.2);
  sor.setInputCloud(std::make_shared<lgs::PointCloud>(filtered));
  lgs::PointCloud inliers;
  sor.filter(inliers);
  if (!sor.ok() || inliers.empty() || inliers.size() > filtered.size()) return 2;
  std::shared_ptr<lgs::Registration> registration = std::make_shared<lgs::NormalDistributionsTransform>();
  auto ndt = std::static_pointer_cast<lgs::NormalDistributionsTransform>(registration);
  ndt->setTransformationEpsilon(0.01);
  ndt->setStepSize(0.1);
  ndt->setResolution(1.0f);
  ndt->setMaximumIterations(64);
  ndt->setNeighborhoodSearchMethod(lgs::DIRECT7);
  registration->setInputTarget(cloud);
  registration->setInputSource(std::make_shared<lgs::PointCloud>(filtered));
  lgs::PointCloud aligned;
  registration->align(aligned);
  std::printf("filtered %zu -> converged %d, iterations %d, fitness %.6f\n", filtered.size(), int(registration->hasConverged()),
              ndt->getFinalNumIteration(), registration->getFitnessScore());
  if (!registration->hasConverged()) return 1;
  // the "GICP" method of the nodes (LSM:73-96): pclomp::GeneralizedIterativeClosestPoint with the scan matcher's setters
  auto gicp = std::make_shared<lgs::GeneralizedIterativeClosestPoint>();
  gicp->setMaxCorrespondenceDistance(2.0);
  gicp->setMaximumIterations(30);
  gicp->setUseReciprocalCorrespondences(false);
  gicp->setMaximumOptimizerIterations(20);
  gicp->setTransformationEpsilon(0.01);
  gicp->setEuclideanFitnessEpsilon(0.01);
  gicp->setCorrespondenceRandomness(20);
  registration = gicp;
  registration->setInputTarget(cloud);
  registration->setInputSource(std::make_shared<lgs::PointCloud>(filtered));
  registration->align(aligned);
  std::printf("GICP (BFGS): converged %d, outer iterations %d, fitness %.6f\n", int(registration->hasConverged()), registration->result().iterations,
              registration->getFitnessScore());
  if (!registration->hasConverged()) return 3;
  // key-frame array: two key frames 1 m apart, sub-map on the device, handed to the registration as a device cloud
  lgs::KeyFrameArray key_frames;
  lgs::Matrix4f pose = lgs::Identity4f();
  const int id0 = key_frames.push(*cloud, pose, 0.0);
  pose[12] = 1.0f;
  const int id1 = key_frames.push(filtered, pose, 1.0);
  const float* map_dev = nullptr;
  int64_t n_map = 0;
  if (id0 != 0 || id1 != 1 || !key_frames.assemble({id1, id0}, 0.0f, &map_dev, &n_map)) return 4;
  if (n_map != static_cast<int64_t>(cloud->size() + filtered.size())) return 5;
  if (lgs_ndt_set_target_dev(ndt->handle(), map_dev, n_map) != LGS_OK) return 6;
  if (key_frames.detectLoop(id1, 100.0, 15.0) != -1) return 7;  // 1 m of path: no loop yet
  std::printf("key-frame array: %zu key frames, sub-map of %lld points resident on the GPU\n", key_frames.size(), static_cast<long long>(n_map));
  return 0;
}
