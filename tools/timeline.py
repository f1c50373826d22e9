#!/usr/bin/env python
"""In-situ GPU timeline of the C3 train step (nsys is not in the image): CUPTI kernel records through
torch.profiler, so kernels of the main and the side stream overlap as they do in the timed loop
(ncu serialises them and runs each one cold).  Prints, per kernel of one step, stream / start offset /
duration, and the idle time of the main stream (the gaps a CUDA graph or PDL could remove).

usage: python tools/timeline.py [--steps 6] [--out gpurun_out/timeline.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--e2e", action="store_true", help="profile bench.py's end-to-end loop (pinned H2D per step, loss D2H)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.json"))
    args = ap.parse_args()
    import svox2_b200 as sv
    from svox2_b200.dist import ShardedTrainer

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    sc = bench.build_scene(args.workload, dev)
    grid = sv.SparseGrid(reso=[2, 2, 2], basis_dim=9, device=dev)
    grid.links = sc["links"]
    grid.density_data = torch.nn.Parameter(sc["density"])
    grid.sh_data = torch.nn.Parameter(sc["sh"])
    grid.capacity = sc["n_rows"]
    grid.accelerate()
    for k, v in bench.RENDER.items():
        setattr(grid.opt, k, v)
    origins, dirs = bench.build_rays(bench.POOL * bench.BATCH, dev)
    with torch.no_grad():
        gt = torch.cat([grid.volume_render(sv.Rays(origins[i:i + 50000], dirs[i:i + 50000]))
                        for i in range(0, origins.size(0), 50000)])
    batches = [(origins[b * bench.BATCH:(b + 1) * bench.BATCH].contiguous(),
                dirs[b * bench.BATCH:(b + 1) * bench.BATCH].contiguous(),
                gt[b * bench.BATCH:(b + 1) * bench.BATCH].contiguous()) for b in range(bench.POOL)]
    trainer = ShardedTrainer(grid, world_size=1, rank=0, **bench.SYN)
    np.random.seed(1234)
    for i in range(args.warmup):
        trainer.step(sv.Rays(*batches[i % bench.POOL][:2]), batches[i % bench.POOL][2], global_batch=bench.BATCH)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    def e2e_loop(steps):    # the loop of bench.py's e2e leg
        host = [tuple(x.cpu().pin_memory() for x in b) for b in batches]
        copy_stream, loss_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream(dev)
        stage = [tuple(torch.empty_like(x, device=dev) for x in host[0]) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ev = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[i & 1])
                for dst, src in zip(stage[i & 1], host[i % bench.POOL]):
                    dst.copy_(src, non_blocking=True)
                ready[i & 1].record(copy_stream)

        for ev in free:
            ev.record(main_stream)
        prefetch(0)
        for i in range(steps):
            if i + 1 < steps:
                prefetch(i + 1)
            main_stream.wait_event(ready[i & 1])
            o, d, g = stage[i & 1]
            trainer.step(sv.Rays(o, d), g, global_batch=bench.BATCH)
            free[i & 1].record(main_stream)
            with torch.cuda.stream(loss_stream):
                loss_stream.wait_event(free[i & 1])
                loss_host[i & 1].copy_(trainer.last_mse, non_blocking=True)
                loss_ev[i & 1].record(loss_stream)
            if i > 0:
                loss_ev[(i - 1) & 1].synchronize()
        torch.cuda.synchronize()

    if args.e2e:
        e2e_loop(4)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        if args.e2e:
            e2e_loop(args.steps)
        else:
            for i in range(args.steps):
                trainer.step(sv.Rays(*batches[i % bench.POOL][:2]), batches[i % bench.POOL][2],
                             global_batch=bench.BATCH)
        torch.cuda.synchronize()
    evs = []
    for e in prof.profiler.kineto_results.events():     # raw CUPTI activity records (stream ids included)
        if e.device_type() == torch.autograd.DeviceType.CUDA:
            evs.append({"name": e.name()[:90], "start_us": e.start_ns() / 1e3, "dur_us": e.duration_ns() / 1e3,
                        "stream": e.device_resource_id()})
    evs.sort(key=lambda x: x["start_us"])
    if not evs:
        print("no CUDA events captured")
        return
    t0 = evs[0]["start_us"]
    for e in evs:
        e["start_us"] -= t0
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(evs, open(args.out, "w"))
    # steps begin at step_clear_kernel; report the last complete one
    starts = [i for i, e in enumerate(evs) if "step_clear" in e["name"]]
    if len(starts) >= 3:
        a, b = starts[-2], starts[-1]
        one = evs[a:b]
        base = one[0]["start_us"]
        main_stream = one[0]["stream"]
        print("one step (%.1f us from clear to the next clear):" % (evs[b]["start_us"] - base))
        busy_end = base
        idle = 0.0
        for e in one:
            tag = "main" if e["stream"] == main_stream else ("s%s" % e["stream"])[:4]
            print("  %-4s +%8.1f us  %7.1f us  %s" % (tag, e["start_us"] - base, e["dur_us"], e["name"][:70]))
            if tag == "main":
                idle += max(0.0, e["start_us"] - busy_end)
                busy_end = max(busy_end, e["start_us"] + e["dur_us"])
        idle += max(0.0, evs[b]["start_us"] - busy_end)
        print("main-stream idle inside the step: %.1f us" % idle)
        per = [(evs[starts[i + 1]]["start_us"] - evs[starts[i]]["start_us"]) for i in range(len(starts) - 1)]
        print("step periods (us):", [round(p, 1) for p in per])


if __name__ == "__main__":
    main()
