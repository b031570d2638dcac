"""Timeline of one CFG+DDIM step (the captured CUDA graph) from CUPTI activity records (torch.profiler).

    python tools/graph_trace.py [--serial-passes] [--steps 3] [--top 25]

Prints, for the replayed graph: wall span per step, per-stream busy time and idle gaps, the time at least one kernel is
running, and per-kernel totals (warm, in-graph durations: complements the cold-cache serialised ncu launch list).
Not a bench: the numbers are taken under a profiler and only used to decide what to optimise.
"""
import argparse
import collections
import os
import re
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from camc2v_b200.config import UNetConfig  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--serial-passes", action="store_true")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--top", type=int, default=30)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--by-shape", action="store_true", help="(with --serial-passes) attribute warm kernel times to call shapes")
    ap.add_argument("--dump", default=None, help="write every kernel record (stream,start_us,dur_us,name) of the last step here")
    args = ap.parse_args()
    device = torch.device("cuda", 0)
    cfg = UNetConfig()
    B = args.batch
    model, sampler, host, cam_host, _ = bench.build_workload(cfg, B, device)
    sampler.concurrent_passes = not args.serial_passes
    cond, uc, static, _ = bench.to_device_conditioning(host, cam_host, device)
    kw = dict(unconditional_guidance_scale=3.5, unconditional_conditioning=uc, guidance_rescale=0.7, fs=static["fs"],
              enable_camera_condition=True, use_cuda_graph=True)
    ts_table = np.flip(sampler.ddim_timesteps).copy()

    def step(x, i):
        ts = torch.full((B,), int(ts_table[i % 25]), device=device, dtype=torch.long)
        return sampler.p_sample_ddim(x, cond, ts, index=24 - (i % 25), **kw)[0]

    x = host["x"].to(device)
    for i in range(4):
        x = step(x, i)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(args.steps):
            x = step(x, i)
            torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower()
          and "memset" not in e.name.lower()]
    recs = sorted(((e.time_range.start, e.time_range.end, getattr(e, "device_resource_id", 0), e.name) for e in ev), key=lambda r: r[0])
    if not recs:
        raise SystemExit("no CUDA kernel records (CUPTI unavailable?)")
    # split into steps by large gaps (host sync between steps)
    steps, cur = [], [recs[0]]
    for r in recs[1:]:
        if r[0] - max(c[1] for c in cur[-8:]) > 200:      # > 200 us of nothing: next step
            steps.append(cur)
            cur = [r]
        else:
            cur.append(r)
    steps.append(cur)
    steps = [s for s in steps if len(s) > 100]
    print(f"{len(steps)} step(s) with {[len(s) for s in steps]} kernels")
    last = steps[-1]
    t0, t1 = min(r[0] for r in last), max(r[1] for r in last)
    print(f"span {1e-3 * (t1 - t0):.3f} ms, sum of kernel durations {1e-3 * sum(r[1] - r[0] for r in last):.3f} ms")
    by_stream = collections.defaultdict(list)
    for r in last:
        by_stream[r[2]].append(r)
    for s, rs in by_stream.items():
        busy = sum(r[1] - r[0] for r in rs)
        gaps = [b[0] - a[1] for a, b in zip(rs, rs[1:])]
        g = np.array(gaps) if gaps else np.zeros(1)
        print(f"  stream {s}: {len(rs)} kernels, busy {1e-3 * busy:.3f} ms, first..last {1e-3 * (rs[-1][1] - rs[0][0]):.3f} ms, "
              f"gaps: sum {1e-3 * g.clip(min=0).sum():.3f} ms median {np.median(g):.2f} us p90 {np.percentile(g, 90):.2f} us max {g.max():.1f} us")
    # union busy
    iv = sorted((r[0], r[1]) for r in last)
    u, ce = 0.0, iv[0][0]
    for a, b in iv:
        if b > ce:
            u += b - max(a, ce)
            ce = b
    print(f"  >=1 kernel running: {1e-3 * u:.3f} ms ({100 * u / (t1 - t0):.1f}% of span)")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in last:
        n = re.sub(r"\(.*", "", r[3]).replace("void ", "").replace("c2v::", "")
        agg[n][0] += 1
        agg[n][1] += r[1] - r[0]
    tot = sum(v[1] for v in agg.values())
    print("kernel,launches,total_us,avg_us,share_pct")
    for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: args.top]:
        print(f"{n},{v[0]},{v[1]:.1f},{v[1] / v[0]:.2f},{100 * v[1] / tot:.2f}")
    if args.by_shape:
        by_shape(model, sampler, cond, uc, static, last)
    if args.dump:
        with open(args.dump, "w") as f:
            f.write("stream,start_us,dur_us,name\n")
            for r in last:
                name = re.sub(r"\(.*", "", r[3])
                f.write(f"{r[2]},{r[0] - t0:.2f},{r[1] - r[0]:.2f},\"{name}\"\n")


def by_shape(model, sampler, cond, uc, static, recs):
    """Replay the two passes eagerly with the C-ABI calls logged, then zip the expected kernel sequence with the
    serial graph's records (same order: the graph was captured from the same Python code)."""
    import ctypes as C
    from camc2v_b200 import _lib
    log = []
    orig = _lib.call

    def spy(name, *a):
        if name == "c2v_gemm":
            d = a[0]._obj
            key = (f"gemm M={d.M} N={d.N} K={d.Cin * d.taps} taps={d.taps} epi={d.epi} res={int(bool(d.residual))} bf16={d.out_bf16} "
                   f"sk={max(1, d.splitk)}")
            fl = 2.0 * d.M * d.N * d.Cin * d.taps
            log.append(("gemm_", key, fl))          # gemm_tc_kernel or gemm_ps_kernel (persistent)
            if d.splitk > 1 and d.ws:
                log.append(("splitk_reduce", f"splitk_reduce M={d.M} N={d.N} sk={d.splitk}", 0.0))
        elif name == "c2v_attention":
            d = a[0]._obj
            key = f"attn bq={d.bq} lq={d.lq} lk={d.lk}+{d.lk2} h={d.heads} kvdiv={d.kv_div} epi={int(bool(d.epi_F))}/{d.epi_d} acc={d.accumulate}"
            log.append(("attn_", key, 4.0 * d.bq * d.lq * (d.lk + d.lk2) * d.heads * 64))
        elif name == "c2v_groupnorm_silu":
            key = f"ns={a[5]} rows={a[6]} C={a[7]}"
            log.append(("gn_stats", "gn_stats " + key, 0.0))
            log.append(("gn_apply", "gn_apply " + key, 0.0))
        elif name == "c2v_layernorm":
            log.append(("layernorm", f"layernorm rows={a[7]} C={a[8]} add={int(bool(a[4]))}", 0.0))
        elif name == "c2v_attention_temporal":
            log.append(("attn_temporal", f"attn_temporal B={a[2]} T={a[3]} HW={a[4]} h={a[5]}", 0.0))
        elif name == "c2v_skinny_linear":
            log.append(("skinny_linear", f"skinny M={a[4]} N={a[5]} K={a[6]}", 0.0))
        else:
            log.append((name.replace("c2v_", "")[:6], name, 0.0))
        return orig(name, *a)

    _lib.call = spy
    try:
        g = sampler._graph
        model.apply_model(g["x"], g["t"], cond, fs=static["fs"], enable_camera_condition=True)
        model.apply_model(g["x"], g["t"], uc, fs=static["fs"], enable_camera_condition=True)
        torch.cuda.synchronize()
    finally:
        _lib.call = orig
    ks = [r for r in recs if "at::native" not in r[3] and "cfg_ddim" not in r[3]]
    print(f"by-shape: {len(log)} expected kernels, {len(ks)} recorded")
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    bad = 0
    for (frag, key, fl), r in zip(log, ks):
        if frag not in r[3]:
            bad += 1
        agg[key][0] += 1
        agg[key][1] += r[1] - r[0]
        agg[key][2] += fl
    print(f"  name mismatches: {bad}")
    print("total_us,launches,avg_us,TFLOP/s,call")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
        print(f"{v[1]:9.1f},{v[0]:4d},{v[1] / v[0]:8.2f},{v[2] / v[1] / 1e6 if v[2] else 0:7.1f},{key}")


if __name__ == "__main__":
    main()
