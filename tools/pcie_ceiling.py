#!/usr/bin/env python
"""Copy-only control of bench.py's end-to-end leg: the same pinned host buffers, the same byte counts per
frame (3840x2160x3 B up, 3840x2160x4 B down), the same number of pipeline threads per rank -- and nothing but one
cudaMemcpyAsync per buffer (torch .copy_(non_blocking=True) of a contiguous pinned tensor) on one stream per thread.
The rate it prints is the ceiling the host I/O path gives the e2e number; no kernel of this repo runs.

  python tools/pcie_ceiling.py                                   # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_ceiling.py

Prints one JSON line on rank 0: frames/s and MP/s summed over ranks (time = max over ranks), GB/s per direction per rank."""
import argparse, json, os, threading, time
import torch

W, H = 3840, 2160


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--mode", default="both", choices=["both", "h2d", "d2h"])
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    frames_host = [torch.randint(0, 255, (H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(a.frames)]
    n_thr = min(a.threads, a.frames)
    host_labels = [torch.empty((H, W), dtype=torch.int32).pin_memory() for _ in range(n_thr)]
    dev_frames = [torch.empty((H, W, 3), dtype=torch.uint8, device=dev) for _ in range(n_thr)]
    dev_labels = [torch.zeros((H, W), dtype=torch.int32, device=dev) for _ in range(n_thr)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_thr)]

    def worker(t, nrep):
        torch.cuda.set_device(local)
        with torch.cuda.stream(streams[t]):
            for _ in range(nrep):
                for i in range(t, a.frames, n_thr):
                    if a.mode != "d2h":
                        dev_frames[t].copy_(frames_host[i], non_blocking=True)
                    if a.mode != "h2d":
                        host_labels[t].copy_(dev_labels[t], non_blocking=True)
                    streams[t].synchronize()        # the e2e leg's getLabels blocks per frame, too

    def run(nrep):
        th = [threading.Thread(target=worker, args=(t, nrep)) for t in range(n_thr)]
        for x in th: x.start()
        for x in th: x.join()

    def barrier():
        if dist is not None: dist.barrier()
        torch.cuda.synchronize()

    run(1); barrier()
    t0 = time.perf_counter(); run(a.reps); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    tt = torch.tensor([dt], device=dev, dtype=torch.float64)
    if dist is not None: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    n = a.frames * a.reps
    if rank == 0:
        print(json.dumps({"tool": "pcie_ceiling", "n_gpus": world, "mode": a.mode, "threads_per_rank": n_thr, "frames_per_rank": n,
                          "frames_per_sec_total": world * n / dt, "MP_s_total": world * n * W * H / 1e6 / dt,
                          "h2d_GBs_per_rank": (0 if a.mode == "d2h" else n * W * H * 3 / dt / 1e9),
                          "d2h_GBs_per_rank": (0 if a.mode == "h2d" else n * W * H * 4 / dt / 1e9)}))
    if dist is not None: dist.destroy_process_group()


if __name__ == "__main__":
    main()
