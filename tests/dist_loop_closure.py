#!/usr/bin/env python
"""Multi-rank leg of the cfg 4 tests (run under torch.distributed.run, one rank per GPU; also works as a single process).

Every rank holds the whole key-frame array (the pose graph is replicated), calls the collective
lgs_batch_align_keyframes_dist with the SAME candidate list and must receive ALL records; rank 0 writes them to --out so
that the pytest wrapper (tests/test_gpu_dist.py) can compare world sizes bit for bit and a sample with the oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tests/dist_loop_closure.py --pairs 4096 --out /tmp/cfg4_w2.npz
"""
import argparse
import ctypes
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def records_bytes(recs):
    return b"".join(ctypes.string_at(ctypes.addressof(r), ctypes.sizeof(r)) for r in recs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4096)
    ap.add_argument("--azimuth", type=int, default=900)
    ap.add_argument("--unique", type=int, default=2)
    ap.add_argument("--workers", type=int, default=0)
    ap.add_argument("--host-arrays", type=int, default=0, help="also run lgs_batch_align_dist on the first N pairs from host arrays")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    import torch
    import torch.distributed as dist
    from lidar_graph_slam_b200 import api, synth
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = synth.loop_keyframes(n_pairs=args.pairs, n_keyframes=41, n_azimuth=args.azimuth, n_unique=args.unique)
    ctx = api.Context(local)
    kf = api.KeyFrameArray(ctx)
    for c, P in zip(d["clouds"], d["poses"]):
        kf.push(c, P)
    comm = api.Comm.from_torch_distributed(local)
    # warm-up: device state of every worker
    kf.batch_align_dist(comm, d["scan_ids"][:8 * world], d["center_ids"][:8 * world], search_key_frame_num=20, n_workers=args.workers)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    recs, info = kf.batch_align_dist(comm, d["scan_ids"], d["center_ids"], search_key_frame_num=20, n_workers=args.workers)
    dt = time.perf_counter() - t0
    blob = records_bytes(recs)
    digest = hashlib.sha256(blob).hexdigest()
    tmax = torch.tensor([dt], dtype=torch.float64, device="cuda:%d" % local)
    digests = [digest]
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        digests = [None] * world
        dist.all_gather_object(digests, digest)
    assert len(set(digests)) == 1, "ranks disagree about the gathered records: %r" % (digests,)
    assert [r.pair_id for r in recs] == list(range(args.pairs))
    host = None
    if args.host_arrays:
        n = args.host_arrays
        scans, submaps, _ = synth.loop_pairs(n_pairs=n, n_keyframes=41, n_azimuth=args.azimuth, n_unique=args.unique)
        sizes = [(len(a), len(b)) for a, b in zip(scans, submaps)]
        mine = set(api.partition_pairs([a + b for a, b in sizes], rank, world))
        # every rank passes only the clouds it owns
        hrecs, hinfo = api.batch_align_dist(comm, [s if i in mine else None for i, s in enumerate(scans)],
                                            [s if i in mine else None for i, s in enumerate(submaps)], sizes=sizes, n_workers=args.workers)
        host = records_bytes(hrecs)
        hd = [hashlib.sha256(host).hexdigest()]
        if world > 1:
            hd = [None] * world
            dist.all_gather_object(hd, hashlib.sha256(host).hexdigest())
        assert len(set(hd)) == 1
    if rank == 0:
        print("world %d: %d pairs in %.3f s = %.1f pairs/s; rank 0 verified %d pairs in %.1f ms, gather %.2f ms (%d bytes), NCCL %d" %
              (world, args.pairs, tmax.item(), args.pairs / tmax.item(), info["n_local"], info["verify_ms"], info["gather_ms"], info["gather_bytes"],
               comm.nccl_version), flush=True)
        if args.out:
            np.savez(args.out, records=np.frombuffer(blob, np.uint8), world=world, seconds=tmax.item(), n_local=info["n_local"],
                     gather_ms=info["gather_ms"], nccl=comm.nccl_version,
                     host_records=np.frombuffer(host, np.uint8) if host is not None else np.zeros(0, np.uint8))
    comm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
