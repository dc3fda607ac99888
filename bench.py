#!/usr/bin/env python
"""bench.py -- IRN influence-path generation throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--users B] [--small]

Workload: synthetic 1M-item catalog, history 200 => window L=201, d=128,
6 layers / 4 heads / ffn 256, reference default initialisers under torch.manual_seed(1234); users in
device tiles of B (default 4096) per GPU.  A *step* is one pass of the hot path over one tile:
gather+PE -> 6 PIM decoder layers (row L-2 only in the last) -> fused catalog scoring + window mask +
arg-max -> window shift, i.e. B user-steps.  value = user-steps/s over all GPUs.

N>1 (torchrun): users are split data-parallel AND the catalog is row-sharded: every step all-gathers the
decoded rows + windows, each rank scores its catalog shard for all users, candidates are all-gathered and
merged (irs_topk_merge).  Per-GPU work is constant as N grows => "scaling": "weak".

One JSON line on stdout (rank 0).  --impl reference times the reference algorithm's CPU port
(oracle/irn_oracle.generate_paths_faithful) on the host cores with a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# dram__bytes_read.sum + dram__bytes_write.sum of one FULL launch at cfg3 / 4096 users from the committed ncu captures
# (profiles/r1_prof_*_summary.csv), or None where no capture of the current kernel exists
TRAFFIC_BYTES = {"scorer": 276.9e6, "attention": 1.889e9, "decoder_chain": 2.568e9, "gather": None}
KERNEL_NAMES = {"attention": "pim_attn_persistent_kernel (tcgen05 PIM attention from operand images)",
                "decoder_chain": "decoder_chain_kernel (fused out_proj+LN1+LN2 -> FFN+LN3 -> next in_proj, tcgen05)",
                "scorer": "score_tc_kernel<0> + rescore_finalize_kernel (fused catalog scorer: one bf16 tcgen05 MMA per K step, rigorous "
                          "rounding-error band, exact fp32 re-score of the candidates)",
                "gather": "embed_gather_v4_kernel (item embedding gather + sqrt(d) + PE)"}


def workload_config(cfg, B, world, small=False):
    return {"workload": "cfg3: IRN generation, 1M-item catalog, L=201 (history 200 + objective), d=128, "
                        "6 layers/4 heads/ffn 256" if not small else "small", "users_per_gpu": B,
            "n_item": cfg["n_item"], "catalog_shards": world, "l2_policy": "inputs > L2 (W 512 MB, E 512 MB)",
            "weights": "reference default init, seed 1234"}

CFG3 = dict(n_item=1_000_000, n_user=100_000, max_len=201, n_layers=6, n_heads=4, emb_dim=128, u_emb_dim=10,
            ffn_dim=256, dropout=0.0, lr1=1e-3)
SMALL = dict(n_item=20_000, n_user=1_000, max_len=41, n_layers=2, n_heads=4, emb_dim=128, u_emb_dim=10,
             ffn_dim=256, dropout=0.0, lr1=1e-3)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def synth_batch(B, cfg, gen, device):
    """[0-free] windows: 200 distinct uniform history ids + a target not in the history (cfg3 i)."""
    L, N = cfg["max_len"], cfg["n_item"]
    H = L - 1
    stride = N // L
    base = torch.arange(L, device=device).unsqueeze(0) * stride
    off = torch.randint(0, stride, (B, L), generator=gen, device=device)
    ids = base + off + 1                                            # distinct by construction
    perm = torch.rand((B, L), generator=gen, device=device).argsort(1)
    seqs = ids.gather(1, perm).contiguous()                         # last column = objective item
    users = torch.randint(0, cfg["n_user"], (B,), generator=gen, device=device)
    return seqs, users


def build_model(cfg, device):
    from types import SimpleNamespace
    import influentialrs_b200 as pkg
    torch.manual_seed(1234)
    c = SimpleNamespace(**cfg)
    net = pkg.InfluentialNet(c)
    net.to(device).eval()
    irn = pkg.IRSNN(c, net, device)
    return pkg, c, net, irn


# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: influentialrs_b200 has no CPU path")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    cfg = dict(SMALL if args.small else CFG3)
    B = args.users
    pkg, c, net, irn = build_model(cfg, device)
    ops = pkg.ops
    stepper = None
    if world > 1:
        from influentialrs_b200.dist import ShardedGenerator
        stepper = ShardedGenerator(irn, rank, world)
    gen = torch.Generator(device=device).manual_seed(1234 + rank)
    seqs, users = synth_batch(B, cfg, gen, device)
    L = cfg["max_len"]
    p = L - 2
    W, beta = net.project.weight, net.project.bias
    paths = torch.zeros((B, args.steps + args.warmup), dtype=torch.float32, device=device)
    temp = seqs.clone()
    def step(i, timed):
        with torch.no_grad():
            if stepper is not None:
                stepper.step(temp, users, paths, i)
                return
            h = net.decoding(temp, users, last_row=p)
            excl = ops.sort_exclusions(temp[:, : p + 1], cfg["n_item"], 1)
            nxt = irn.next_items(h, excl)
            ops.window_shift(temp, nxt, paths, i)

    for i in range(args.warmup):
        step(i, False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.launch_count_reset()
    ops._timer = {}                 # per-kernel-class CUDA events on the launching stream, inside the timed region
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for i in range(args.steps):
        step(args.warmup + i, True)
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = t0.elapsed_time(t1)
    launches = ops.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    timer, ops._timer = ops._timer, None
    # (ms/launch, launches/step); a record is one or more (start, stop) event pairs (the two-phase sharded scorer has two)
    kms = {k: (sum(sum(x[i].elapsed_time(x[i + 1]) for i in range(0, len(x), 2)) for x in v) / len(v), len(v) / args.steps)
           for k, v in timer.items()}
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # ---- end to end through the public API with HOST buffers (H2D of the batch + D2H of the paths inside)
    P_e2e = args.e2e_path_len
    seqs_h = seqs.cpu().pin_memory()
    users_h = users.cpu().pin_memory()
    tg_h = seqs_h[:, -1].clone().pin_memory()

    def e2e_call():
        s = seqs_h.to(device, non_blocking=True)
        u = users_h.to(device, non_blocking=True)
        t = tg_h.to(device, non_blocking=True)
        if stepper is not None:
            return stepper.get_seq_in_batch(s, u, t, max_path_len=P_e2e)
        return irn.get_seq_in_batch(s, u, t, max_path_len=P_e2e, gap_len=0)
    for _ in range(args.e2e_warm):
        e2e_call()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    w0 = time.perf_counter()
    out = e2e_call()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    if world > 1:
        t = torch.tensor([e2e_s], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = (seqs_h.numel() + users_h.numel() + tg_h.numel()) * 8
    d2h = out[0].size * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    value = B * world * args.steps / (ms / 1e3)
    step_ms = ms / args.steps
    L, d, ffn, n_layers = cfg["max_len"], cfg["emb_dim"], cfg["ffn_dim"], cfg["n_layers"]
    n_shard = cfg["n_item"] // world
    # ALGORITHMIC work per launch (SURVEY.md section 8d / DESIGN.md section 4); B users per launch
    alg = {
        # averages over the n_layers launches of a step: the last layer's attention is the one-row kernel (4*L*d FLOP per
        # user), the first chain launch is the in_proj-only mode (x read, q,k,v written: 4*L*d*4 B per user)
        "attention": ("tensor", ((n_layers - 1) * 4.0 * L * L * d + 4.0 * L * d) / n_layers * B,
                      "4*L^2*d FLOP per user per full layer (QK^T + PV over the full window; the kernel issues 3x that in bf16 MMAs "
                      "minus the causally invisible key blocks), 4*L*d for the one-row last layer; average over the step's launches"),
        "decoder_chain": ("hbm", ((n_layers - 1) * 6.0 + 4.0) / n_layers * L * d * 4 * B,
                          "6*L*d*4 B per user per full launch (attn + x read, x' + q,k,v written, fp32-equivalent), 4*L*d*4 for the "
                          "first layer's in_proj-only launch; average over the step's launches"),
        "scorer": ("tensor", 2.0 * d * n_shard * (B * world), "2*d*N FLOP per user-step, issued once in bf16 (hi*hi); fp32-faithful winners "
                   "come from the rigorous error band + exact re-scoring (the three-MMA variant issues 3x this)"),
        "gather": ("hbm", float(L * (8 + 4 * d + 4 * d)) * B, "L*(8 + 4d + 4d) B per user-step: id + table row read + row written"),
    }
    kernels = {}
    for name, (bound, work, note) in alg.items():
        if name not in kms:
            continue
        ms_l, per_step = kms[name]
        peak = pk["tf_sus"] if bound == "tensor" else pk["hbm"]
        ach = work / (ms_l / 1e3) / (1e12 if bound == "tensor" else 1e9)
        kernels[name] = {"bound": bound, "achieved": ach, "peak": peak, "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                         "frac": ach / peak, "traffic": TRAFFIC_BYTES.get(name), "ms_per_launch": ms_l,
                         "launches_per_step": per_step, "share_of_step": ms_l * per_step / step_ms, "algorithmic": note}
    dominant = max(kernels, key=lambda k: kernels[k]["share_of_step"]) if kernels else None
    roof = dict(kernels[dominant], kernel=KERNEL_NAMES[dominant],
                peak_source=pk["source"] + (" bf16 sustained" if kernels[dominant]["bound"] == "tensor" else " HBM copy")
                + " (kernel timed inside the step)") if dominant else None
    line = {
        "metric": "IRN influence-path generation throughput @1M items" if not args.small else "IRN generation (small)",
        "value": value, "unit": "user-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, B, world, args.small),
        "e2e": {"value": B * world * P_e2e / e2e_s, "unit": "user-steps/s", "h2d_bytes_per_step": h2d / P_e2e,
                "d2h_bytes_per_step": d2h / P_e2e, "path_len": P_e2e, "api": "IRSNN.get_seq_in_batch"},
        "gpu_launches": int(launches),
        "roofline": roof,
        "roofline_kernels": {KERNEL_NAMES[k]: v for k, v in kernels.items()},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg, args)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
def cpu_sample(cfg, users, steps, threads):
    """The reference algorithm's CPU port on a bounded sample: materialise [B,L,N] logits, softmax,
    per-sample top-100 + window filter (oracle.generate_paths_faithful)."""
    from oracle import irn_oracle as O       # bench.py's cpu_baseline / --impl reference legs only
    torch.set_num_threads(threads)
    sd = O.synth_irn_state(cfg["n_item"], min(cfg["n_user"], 1000), cfg["max_len"], cfg["emb_dim"], cfg["n_layers"],
                           cfg["ffn_dim"], seed=1234)
    g = torch.Generator().manual_seed(1234)
    seqs, us = synth_batch(users, cfg, g, torch.device("cpu"))
    us = us % 1000
    t0 = time.perf_counter()
    with torch.no_grad():
        O.generate_paths_faithful(sd, seqs, us, seqs[:, -1], cfg["n_heads"], max_path_len=steps)
    return time.perf_counter() - t0


def cpu_baseline(cfg, args):
    """Bounded sample (~10-20 s of CPU work): the reference batches a few users at a time because it
    materialises [B,L,N] logits twice (1.6 GB per user at cfg3)."""
    threads = os.cpu_count() or 1
    users, steps = (4, 2) if cfg["n_item"] >= 500_000 else (32, 2)
    reps, dt = 0, 0.0
    while dt < 10.0 and reps < 12:
        dt += cpu_sample(cfg, users, steps, threads)
        reps += 1
    return {"value": reps * users * steps / dt, "unit": "user-steps/s", "cores": threads, "kind": "port",
            "sample": f"{reps} batches of {users} users x {steps} path steps of the same workload (materialised "
                      f"[B,L,N] logits + softmax + top-100 + window filter), torch CPU fp32, {dt:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = dict(SMALL if args.small else CFG3)
    threads = os.cpu_count() or 1
    users, steps = (4, 2) if cfg["n_item"] >= 500_000 else (32, 2)
    for _ in range(min(args.warmup, 1)):
        cpu_sample(cfg, users, steps, threads)
    n = max(1, min(args.steps, 3))
    t = [cpu_sample(cfg, users, steps, threads) for _ in range(n)]
    dt = sum(t) / len(t)
    v = users * steps / dt
    cb = {"value": v, "unit": "user-steps/s", "cores": threads, "kind": "port",
          "sample": f"{users} users x {steps} path step(s) per bench step, {n} timed steps, torch CPU fp32"}
    print(json.dumps({
        "impl": "reference", "metric": "IRN influence-path generation throughput @1M items", "value": v,
        "unit": "user-steps/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": n, "warmup": min(args.warmup, 1),
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": dict(workload_config(cfg, args.users, int(os.environ.get("WORLD_SIZE", "1")), args.small),
                                            cpu_sample=f"{users} users x {steps} path steps per bench step"),
        "cpu_baseline": cb, "e2e": {"value": v, "unit": "user-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference algorithm restated on CPU (the reference is PyTorch-only and cannot travel to the GPU box)"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--users", type=int, default=4096, help="users per GPU per step")
    ap.add_argument("--small", action="store_true", help="tiny catalog (debug)")
    ap.add_argument("--e2e-path-len", type=int, default=20)
    ap.add_argument("--e2e-warm", type=int, default=1, help="untimed end-to-end calls before the timed one")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
