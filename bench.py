#!/usr/bin/env python3
"""bench.py — AR-CVAE train molecules/s on B200 (BASELINE.json metric), beside the CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload train|sample]

One "step" = one full default-config training step (encoder fwd + BPTT, decoder fwd + bwd, fused loss, two Adam
updates) on one synthetic SELFIES-shaped batch of B=4096 x T=128 per GPU (BASELINE configs[1]; N>1: data parallel,
global batch 4096*N, configs[2] at N=8 -> "scaling": "weak").
  value : molecules/s, inputs already resident in HBM, CUDA events, barrier + synchronize on both sides, max over ranks
  e2e   : the same step through the public API from pinned HOST buffers: H2D copy of tokens / conditions / eps and a
          D2H read of the loss inside the timed region
  roofline : the dominant kernel category, timed with CUDA events on the launch stream in a second, instrumented pass
  cpu_baseline : the CPU restatement of the reference step (oracle/, torch fp32, all host threads) on a bounded sample
--impl reference times that CPU restatement alone (MLX is not installable here: "kind": "port").
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = dict(vocab_size=80, embedding_dim=128, hidden_dim=256, latent_dim=128, num_conditions=1, num_layers=2)
HYPER = dict(beta=0.05, lambda_prop=0.1, lambda_collapse=0.001, free_bits=1.0, lambda_mi=0.01)
LR = 2e-4
B_PER_GPU, T = 4096, 128
CPU_SAMPLE_B = 256            # molecules per CPU step (the CPU figure may be taken at B >= 256 of the same workload)
CPU_BUDGET_S = 240.0          # the reference arm shrinks its per-step sample if K+W steps would not fit this
# algorithmic FLOP of the reference formulation
FLOP_FWD_PER_MOLECULE = 128 * (1_835_008 + 829_440) + 786_944          # 341,836,288 at T=128
FLOP_STEP_PER_MOLECULE = 3 * FLOP_FWD_PER_MOLECULE                     # fwd + bwd = 3x
LOSS_BYTES_PER_MOLECULE = T * 80 * 4 * 2 + T * 4                       # logits read + dlogits write + tokens


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, sampled in-process through NVML every 50 ms
    (B200_PROFILING.md recipe).  An `nvidia-smi -lms` child process was measurably perturbing the step (a query with
    power and event-reason fields every 100 ms cost ~15 % of a 20 ms step), so the bench polls NVML from a thread."""

    def __init__(self, gpu_index):
        import threading
        self.idx = gpu_index
        self.rows = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.err = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                try:
                    phys = int(vis.split(",")[gpu_index])
                except Exception:
                    phys = gpu_index
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv = None
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, reasons))
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                return
            self.stop_flag.wait(0.05)

    def start(self):
        if self.nv is None:
            return
        import threading
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.nv is None or self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + str(self.err)]}
        time.sleep(0.06)
        self.stop_flag.set()
        self.thread.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        sm = sorted(r[0] for r in self.rows)
        reasons = set()
        for _, mask in self.rows:
            for n, bit in names.items():
                if mask & bit:
                    reasons.add(n)
        return {"sm_mhz": float(sm[len(sm) // 2]) if sm else None, "sm_max_mhz": float(self.max_sm),
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml"}


def cpu_step_rate(steps, warmup, B=CPU_SAMPLE_B, budget_s=None):
    """molecules/s of the CPU restatement of the reference step (trainer.py:292-333: value_and_grad, no-op clip, two
    Adam updates) on this host's cores, torch fp32, all threads.  Every step processes B molecules of the benched
    workload (same dims, same T).  With `budget_s`, B is halved (not below 32) until warmup + steps fit the budget,
    judged from the first step.  Returns (molecules/s, ms per step, threads, B used)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import arcvae_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.Config(**DIMS)
    while True:
        params = O.init_params(cfg, seed=67, dtype=torch.float32)
        state = O.adam_init(params)
        x, cond, eps, tf_mask = O.synthetic_batch(B, T, cfg, seed=67, tf_ratio=0.9)
        xt, ct, et = torch.as_tensor(x), torch.as_tensor(cond), torch.as_tensor(eps)
        times, shrink = [], False
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            vals, grads, params, state = O.train_step(params, state, xt, ct, cfg.num_layers, et, tf_mask, LR,
                                                      target_mi=4.85, **HYPER)
            float(vals["total_loss"])
            dt = time.perf_counter() - t0
            if i == 0 and budget_s is not None and dt * (warmup + steps) > budget_s and B > 32:
                shrink = True
                break
            if i >= warmup:
                times.append(dt)
        if not shrink:
            break
        B //= 2
    ms = 1e3 * sum(times) / len(times)
    return B / (ms / 1e3), ms, torch.get_num_threads(), B


SAMPLE_B_PER_GPU, SAMPLE_T = 125_000, 128       # BASELINE configs[4]: 1M molecules over 8 GPUs, max length 128


def bench_sampler(M, vae, steps=3, warmup=1, B=SAMPLE_B_PER_GPU, T=SAMPLE_T, cpu=True):
    """Sampled molecules/s of generate_with_temperature (greedy = the reference's argmax path, multinomial = Philox
    categorical) on B TPSA conditions swept over a z-scored grid; early stopping off so every row runs T steps.
    resident: conditions already in HBM, tokens left in HBM; e2e: conditions from pinned host memory, tokens read back."""
    import numpy as np
    import torch
    sampler = M.MLXAutoregressiveDecoderSampling(**DIMS, decoder=vae.decoder)
    grid = np.linspace(-2.5, 2.5, B, dtype=np.float32).reshape(B, 1)
    hc = torch.as_tensor(grid).pin_memory()
    dc = hc.cuda()
    out = {"workload": f"TPSA-conditioned sampling, B={B} molecules x max_length {T} per GPU (configs[4] shard), "
                       "early stopping off", "unit": "molecules/s"}
    for name, multi in (("greedy", False), ("multinomial", True)):
        def run(c):
            return sampler.generate_with_temperature(None, c, max_length=T, temperature=1.0, early_stopping=False,
                                                     multinomial=multi, seed=67)
        for _ in range(warmup):
            run(dc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = M._lib.launch_count()
        e0.record()
        for _ in range(steps):
            toks = run(dc)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        launches = (M._lib.launch_count() - l0) // steps
        host = torch.empty((B, T), dtype=torch.int32).pin_memory()       # result buffer of the e2e leg (pinned: D2H at PCIe speed)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            toks = run(hc.cuda(non_blocking=True))
            host[:, : toks.shape[1]].copy_(toks, non_blocking=True)
            torch.cuda.synchronize()
        ms_e2e = 1e3 * (time.perf_counter() - t0) / steps
        out[name] = {"value": B / (ms * 1e-3), "ms": ms, "e2e": B / (ms_e2e * 1e-3), "e2e_ms": ms_e2e,
                     "gpu_launches": launches, "us_per_step_per_128_rows": 1e3 * ms / T / ((B + 127) // 128) * 148,
                     "d2h_bytes": int(host.numel()) * 4, "h2d_bytes": B * 4}
    if cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import arcvae_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        cfg = O.Config(**DIMS)
        params = O.init_params(cfg, seed=67, dtype=torch.float32)
        cb = 256
        ct = torch.as_tensor(grid[:cb])
        t0 = time.perf_counter()
        O.generate_with_temperature(params["decoder"], torch.zeros(cb, cfg.latent_dim), ct, cfg.num_layers, max_length=T,
                                    temperature=1.0, early_stopping=False)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": cb / dt, "unit": "molecules/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"{cb} molecules x {T} steps, greedy, oracle torch-CPU fp32"}
    return out


def workload_config(args, B, world):
    """The `config` object of the JSON line — identical in both arms (the driver compares them)."""
    if args.config == "scaled":
        wl = f"scaled AR-CVAE (V80 E128 H1024 L256 C1 NL3) full train step, B={B} x T={T} per GPU (configs[3]); global batch {B * world}"
    else:
        wl = f"default AR-CVAE (V80 E128 H256 L128 C1 NL2) full train step, B={B} x T={T} per GPU (configs[1]); global batch {B * world}"
    return {"workload": wl, "parallelism": f"dp{world}", "teacher_forcing": 0.9,
            "l2": "no explicit flush: every step streams >6 GB of activations (>> 126 MB L2)"}


def run_reference(args):
    """--impl reference: the reference's CPU path for the same metric and config.  MLX is not installable in this image,
    so this times oracle/arcvae_oracle.py (pinned to the unmodified reference sources by tests/test_ref_pin.py), torch
    fp32 on all host threads; exactly --steps timed steps after --warmup, each a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, args.gpus)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    rate, ms, cores, b_used = cpu_step_rate(steps, warmup, budget_s=CPU_BUDGET_S)
    sample = (f"{b_used} molecules per step (a slice of the B={args.batch} x T={T} batch, same dims and length), "
              f"{steps} timed steps after {warmup} warm-up, torch fp32 eager on {cores} threads")
    line = {"impl": "reference", "metric": "AR-CVAE train molecules/s", "value": rate, "unit": "molecules/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.batch, world),
            "cpu_baseline": {"value": rate, "unit": "molecules/s", "cores": cores, "kind": "port", "sample": sample,
                             "note": "MLX is not installable in this image; this is oracle/arcvae_oracle.py, the torch-CPU "
                                     "restatement pinned to the unmodified reference sources (tests/test_ref_pin.py). "
                                     "molecules/s of a CPU step is flat in B for B >= 64 (time is linear in B)"},
            "e2e": {"value": rate, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--precision", default=os.environ.get("ARCVAE_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sampler", action="store_true", help="skip the sampled-molecules/s leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the fp32-mode and C1 (batch 64 through the dataset) legs")
    ap.add_argument("--workload", default="train", choices=["train", "sample"],
                    help="sample = BASELINE configs[4]: TPSA-conditioned sampling of 1M molecules sharded over the GPUs "
                         "(125k per GPU, max length 128, greedy and multinomial); prints its own JSON line")
    ap.add_argument("--config", default="default", choices=["default", "scaled"],
                    help="scaled = BASELINE configs[3]: hidden 1024, latent 256, 3 layers, length 256 (not the headline; "
                         "per-step tensor-core recurrence, B defaults to 1024)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    global T, FLOP_FWD_PER_MOLECULE, FLOP_STEP_PER_MOLECULE, LOSS_BYTES_PER_MOLECULE
    if args.config == "scaled":
        DIMS.update(hidden_dim=1024, latent_dim=256, num_layers=3)
        T = 256
        if args.batch == B_PER_GPU:
            args.batch = 1024
        args.no_sampler = True
        args.no_cpu = True
        # encoder 42,991,616 + decoder 17,997,824 FLOP per token, head 10,487,808 per molecule
        FLOP_FWD_PER_MOLECULE = T * (42_991_616 + 17_997_824) + 10_487_808
        FLOP_STEP_PER_MOLECULE = 3 * FLOP_FWD_PER_MOLECULE
        LOSS_BYTES_PER_MOLECULE = T * 80 * 4 * 2 + T * 4

    import numpy as np
    import torch
    import torch.distributed as dist
    import mlx_vae_b200 as M
    from mlx_vae_b200.data import synthetic_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    # every rank gets its own shard of the synthetic global batch; the teacher-forcing coins are global (one per position)
    x, cond, eps, _ = synthetic_batch(B, T, seed=67 + rank, tf_ratio=0.9)
    _, _, _, tf_mask = synthetic_batch(8, T, seed=67, tf_ratio=0.9)
    vae = M.ARCVAE(**DIMS, seed=67, precision=args.precision)
    trainer = M.ARCVAETrainerWithLoss(vae.encoder, vae.decoder, None, None, learning_rate=LR, batch_size=B,
                                      lambda_prop=HYPER["lambda_prop"], lambda_collapse=HYPER["lambda_collapse"],
                                      free_bits=HYPER["free_bits"], lambda_mi=HYPER["lambda_mi"])
    beta, tf = HYPER["beta"], 0.9
    dx, dc, de = (torch.as_tensor(a).cuda() for a in (x, cond, eps))
    hx, hc, he = (torch.as_tensor(a).pin_memory() for a in (x, cond, eps))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return trainer.train_step(dx, dc, beta, tf, eps=de, tf_mask=tf_mask)

    def step_e2e():
        gx = hx.cuda(non_blocking=True); gc = hc.cuda(non_blocking=True); ge = he.cuda(non_blocking=True)
        d = trainer.train_step(gx, gc, beta, tf, eps=ge, tf_mask=tf_mask)
        return float(d["total_loss"])                                   # D2H read of the step's result

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = M._lib.launch_count()
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), M._lib.launch_count() - l0, out

    if args.workload == "sample":
        # every rank samples its own shard (no communication: replicas); whole-job rate = sum of shards / max time
        clk = ClockSampler(local) if rank == 0 else None
        if clk:
            clk.start()
        barrier()
        res = bench_sampler(M, vae, steps=max(1, min(args.steps, 5)), warmup=1, cpu=(rank == 0 and not args.no_cpu))
        barrier()
        clocks = clk.stop() if clk else None
        out = {}
        for name in ("greedy", "multinomial"):
            t = torch.tensor([res[name]["ms"], res[name]["e2e_ms"]], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[name] = (float(t[0]), float(t[1]))
        if rank == 0:
            n = SAMPLE_B_PER_GPU * world
            line = {"metric": "AR-CVAE sampled molecules/s", "value": n / (out["greedy"][0] * 1e-3), "unit": "molecules/s",
                    "n_gpus": world, "steps": max(1, min(args.steps, 5)), "warmup": 1, "ms_per_step": out["greedy"][0],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                    "config": {"workload": f"TPSA-conditioned greedy sampling, {SAMPLE_B_PER_GPU} molecules x max_length "
                                           f"{SAMPLE_T} per GPU (configs[4]: {n} molecules over {world} GPUs), early stopping off",
                               "parallelism": f"replicas x{world}"},
                    "e2e": {"value": n / (out["greedy"][1] * 1e-3), "unit": "molecules/s",
                            "h2d_bytes_per_step": res["greedy"]["h2d_bytes"], "d2h_bytes_per_step": res["greedy"]["d2h_bytes"],
                            "ms_per_step": out["greedy"][1]},
                    "multinomial": {"value": n / (out["multinomial"][0] * 1e-3), "e2e": n / (out["multinomial"][1] * 1e-3)},
                    "gpu_launches": res["greedy"]["gpu_launches"], "clocks": clocks,
                    "roofline": {"kernel": "sampler_fused_kernel", "bound": "mufu", "achieved": res["greedy"]["us_per_step_per_128_rows"],
                                 "peak": 8.3, "unit": "us per step per 128-row tile (lower is better)",
                                 "frac": 8.3 / res["greedy"]["us_per_step_per_128_rows"], "traffic": None,
                                 "note": "floor = 8 MUFU.TANH per hidden unit and step at 16/clk/SM; not an HBM or tensor bound"},
                    "cpu_baseline": res.get("cpu_baseline")}
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches, last = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    trainer.check_device_error()    # the cluster kernels use bounded waits: a protocol time-out raises here instead of going unnoticed
    for _ in range(2):
        step_e2e()
    ms_e2e, _, loss_val = timed(step_e2e, args.steps)

    # instrumented pass: per-category device time (CUDA events on the launch stream around every launch scope) and the
    # tensor-core FLOP the library actually issued.  The timed region above runs the two-stream schedule (decoder chain
    # under the encoder's cluster recurrence); a kernel's own duration can only be read without a concurrent stream, so
    # this pass runs the SERIAL order (ARCVAE_ONE_STREAM) — `serial_ms_per_step` is its step time
    os.environ["ARCVAE_ONE_STREAM"] = "1"
    torch.cuda.synchronize()
    torch.cuda.empty_cache()          # the serial order allocates from another stream's pool: settle the caching allocator
    for _ in range(3):                # outside the timed steps (a cudaMalloc / cudaFree stalls the device for tens of ms)
        step_resident()
    ms_serial, _, _ = timed(step_resident, min(args.steps, 5))
    M._lib.timing_enable(True)
    M._lib.timing_read()
    nprof = min(args.steps, 3)
    f0 = {k: M._lib.flop_count(k) for k in ("gemm_tc", "recurrence")}
    for _ in range(nprof):
        step_resident()
    cats = M._lib.timing_read()
    executed = {k: (M._lib.flop_count(k) - f0[k]) / nprof for k in f0}
    M._lib.timing_enable(False)
    os.environ.pop("ARCVAE_ONE_STREAM", None)

    # configs[2] as written: global batch 32768 at EVERY N (the headline line keeps 4096 per GPU = weak scaling)
    configs2 = None
    if args.config == "default" and world in (2, 4) and B == B_PER_GPU:
        Bc = 32768 // world
        cx, cc, ce, _ = synthetic_batch(Bc, T, seed=167 + rank, tf_ratio=0.9)
        gx, gc, ge = (torch.as_tensor(a).cuda() for a in (cx, cc, ce))
        for _ in range(3):
            trainer.train_step(gx, gc, beta, tf, eps=ge, tf_mask=tf_mask)
        ms_c2, _, _ = timed(lambda: trainer.train_step(gx, gc, beta, tf, eps=ge, tf_mask=tf_mask), min(args.steps, 5))
        configs2 = {"global_batch": 32768, "per_gpu_batch": Bc, "ms_per_step": ms_c2, "value": 32768 / (ms_c2 * 1e-3),
                    "unit": "molecules/s", "note": "BASELINE configs[2] (strong-scaling reading: fixed global batch)"}
        del gx, gc, ge

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    per_step = {k: v[0] / nprof for k, v in cats.items() if v[1] > 0}
    launches_cat = {k: v[1] / nprof for k, v in cats.items() if v[1] > 0}
    H, NL = DIMS["hidden_dim"], DIMS["num_layers"]
    R = B * T
    # ncu --set full DRAM bytes per launch of the same command (profiles/ncu_traffic.json, written from the committed
    # capture by profiles/scripts/summarize_ncu.py --traffic); only valid for the default B x T
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and B == B_PER_GPU and args.config == "default":
        traffic = json.load(open(tpath))
    roofs = {}
    if "recurrence" in per_step:
        # encoder LSTM recurrence: per layer T steps of [B,H]x[H,4H] forward and [B,4H]x[4H,H] backward
        flop = NL * 2 * (2.0 * 4 * H * H) * R
        t_ms = per_step["recurrence"]
        n = max(1.0, launches_cat["recurrence"])
        ach = flop / (t_ms * 1e-3) / 1e12
        tape = NL * R * (6 * H * 2 * 2 + H * 2 + 4 * H * 2 + H * 4)   # coefficient tape w+r (6 bf16 per unit), h w, dA w, P/dX r
        roofs["recurrence"] = {"kernel": "lstm_fwd3_kernel+lstm_bwd3_kernel" if H == 256 else
                                         ("gemm_tc_kernel<LSTM-step epilogue> per timestep + k_lstm_cell_bwd_b / split-K gemm_tc_kernel (W_hh L2-resident)"
                                          if H % 64 == 0 else "per-step gemm_tc_kernel + k_lstm_cell_*"),
                               "bound": "tensor", "achieved": ach,
                               "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"],
                               "traffic": traffic.get("recurrence"), "peak_source": pk["src"] + " (sustained bf16)",
                               "launches_per_step": n, "avg_launch_ms": t_ms / n, "ms_per_step": t_ms,
                               "share_of_step": t_ms / ms_serial, "algorithmic_flop_per_launch": flop / n,
                               "executed_flop_per_step": executed.get("recurrence"),
                               "hbm_tape_gbs": tape / (t_ms * 1e-3) / 1e9, "hbm_tape_frac": tape / (t_ms * 1e-3) / 1e9 / pk["hbm"],
                               # cluster kernels (H = 256): per step and 128-row tile the tensor core reads the h / dA operand
                               # tile (64 KB) and the resident W_hh slice (128 KB) from shared memory in each of the 4 CTAs;
                               # the exchange moves 3 x 16 KB per CTA through L2 as flag-in-data vectors (forward: h_t,
                               # backward: bf16 partial sums of d h)
                               "smem_operand_gbs": (NL * 2 * T * ((B + 127) // 128) * 4 * 192 * 1024) / (t_ms * 1e-3) / 1e9 if H == 256 else None,
                               "cluster_exchange_gbs": (NL * 2 * T * ((B + 127) // 128) * 4 * 48 * 1024) / (t_ms * 1e-3) / 1e9 if H == 256 else None,
                               "us_per_timestep": 1e3 * t_ms / (NL * 2 * T),
                               "note": "latency-bound by construction (T sequential exchanges between the 4 CTAs of a cluster); algorithmic FLOP = "
                                       "2*4H*H per row-step, forward + d h backward, all layers"}
    if "gemm_tc" in per_step:
        # time-parallel contractions: everything of SURVEY 8d's 3 x 341.8 MFLOP/molecule that is not the recurrence.
        # `achieved` / `frac` use the ALGORITHMIC FLOP of the reference formulation (tier contract); `executed_*` use what the
        # tcgen05 GEMMs really issued (layer-0 projections are table gathers, the decoder's forget gate is never computed)
        flop = FLOP_STEP_PER_MOLECULE * B - NL * 2 * (2.0 * 4 * H * H) * R
        t_ms = per_step["gemm_tc"]
        n = max(1.0, launches_cat.get("gemm_tc", 0))
        ach = flop / (t_ms * 1e-3) / 1e12
        ex = executed.get("gemm_tc", 0.0)
        roofs["gemm_tc"] = {"kernel": "gemm_tc_kernel", "bound": "tensor", "achieved": ach,
                            "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"],
                            "executed_flop_per_step": ex, "executed_tflops": ex / (t_ms * 1e-3) / 1e12,
                            "executed_frac": ex / (t_ms * 1e-3) / 1e12 / pk["bf16_sustained"],
                            "traffic": traffic.get("gemm_tc"), "peak_source": pk["src"] + " (sustained bf16)",
                            "launches_per_step": n, "avg_launch_ms": t_ms / n, "ms_per_step": t_ms, "share_of_step": t_ms / ms_serial,
                            "note": "K <= 768 projections are HBM-bound on B200 (2K FLOP per 2-byte output element, machine "
                                    "balance ~210 FLOP/B); `frac` = algorithmic FLOP of the reference formulation (incl. work "
                                    "this design does not execute), `executed_frac` = issued tensor FLOP; gemm_f32_kernel "
                                    "time is NOT folded in"}
    # the fused loss kernel, timed ALONE on the benched shapes (logits [T,B,V] fp32 in, d logits in place): on the bf16
    # training path the cross-entropy now runs in the fc_out GEMM epilogue and k_loss_fused only handles the [B,L] latent
    # terms (per_category_ms["loss"]); the full kernel still serves complete_vae_loss (evaluation) and the fp32 mode
    try:
        from mlx_vae_b200.losses._fused import fused_loss, make_hyper
        lg = torch.randn((T, B, DIMS["vocab_size"]), device="cuda").transpose(0, 1)
        mu_t = torch.tanh(torch.randn(B, DIMS["latent_dim"], device="cuda"))
        lv_t = torch.tanh(torch.randn(B, DIMS["latent_dim"], device="cuda")) - 1.0
        hp = make_hyper(HYPER["beta"], HYPER["lambda_prop"], HYPER["lambda_collapse"], HYPER["free_bits"], HYPER["lambda_mi"])
        for _ in range(2):
            fused_loss(lg, dx, mu_t, lv_t, hp, eps=de, inplace_dlogits=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fused_loss(lg, dx, mu_t, lv_t, hp, eps=de, inplace_dlogits=True)
        e1.record()
        torch.cuda.synchronize()
        t_ms = e0.elapsed_time(e1) / 5
        by = LOSS_BYTES_PER_MOLECULE * B
        ach = by / (t_ms * 1e-3) / 1e9
        roofs["loss"] = {"kernel": "k_loss_fused", "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": ach / pk["hbm"], "traffic": traffic.get("loss"), "peak_source": pk["src"],
                         "avg_launch_ms": t_ms, "share_of_step": per_step.get("loss", 0.0) / ms_serial,
                         "note": "timed alone on the benched shapes; algorithmic bytes = logits read + dlogits write + tokens "
                                 "(SURVEY 8d: 82.4 KB/molecule).  On the bf16 training path CE + d logits run in the fc_out "
                                 "epilogue (gemm_tc_kernel<2>) and this kernel only sees the latent terms"}
        del lg
    except Exception as e:  # noqa: BLE001
        roofs["loss"] = {"error": repr(e)}
    # headline roofline = the category that takes the larger share of the step, nothing folded into either
    dom = max(("recurrence", "gemm_tc"), key=lambda k: roofs[k]["ms_per_step"] if k in roofs else -1.0) if roofs else None
    roof = roofs.get(dom)
    whole = {"algorithmic_tflops": FLOP_STEP_PER_MOLECULE * B / (ms * 1e-3) / 1e12,
             "algorithmic_frac_of_bf16_peak": FLOP_STEP_PER_MOLECULE * B / (ms * 1e-3) / 1e12 / pk["bf16_sustained"],
             "executed_tflops": sum(executed.values()) / (ms * 1e-3) / 1e12,
             "executed_frac_of_bf16_peak": sum(executed.values()) / (ms * 1e-3) / 1e12 / pk["bf16_sustained"]}
    extra = {"per_category_ms": per_step, "rooflines": roofs, "whole_step": whole,
             "schedule": {"timed_region": "two streams: decoder forward + backward (independent of the encoder in the reference "
                                          "mode, F1) run under the encoder's cluster recurrence, which holds 128 of 148 SMs",
                          "ms_per_step": ms, "serial_ms_per_step": ms_serial,
                          "note": "per_category_ms / rooflines / share_of_step come from the serial-order pass"}}
    if configs2 is not None:
        extra["configs2_global_32768"] = configs2
    if "achieved" in roofs.get("loss", {}):
        extra["loss_kernel_gbs"] = roofs["loss"]["achieved"]
        extra["loss_kernel_frac_of_hbm_peak"] = roofs["loss"]["frac"]

    # fp32 mode (north_star's 1e-3 path) on the same workload, and C1 = configs[0]: batch 64 through MoleculeDataset built
    # from a ChEMBL-schema stand-in (the bundled JSON is absent from the reference mount)
    if args.config == "default" and world == 1 and not args.no_extras:
        try:
            extra["fp32_mode"] = bench_fp32(M, B, dx, dc, de, tf_mask, beta, tf)
        except Exception as e:  # noqa: BLE001
            extra["fp32_mode"] = {"error": repr(e)}
        try:
            extra["c1_batch64_dataset"] = bench_c1(M, args.precision)
        except Exception as e:  # noqa: BLE001
            extra["c1_batch64_dataset"] = {"error": repr(e)}

    cpu = None
    if not args.no_cpu:
        rate, cms, cores, b_used = cpu_step_rate(steps=4, warmup=1)
        cpu = {"value": rate, "unit": "molecules/s", "cores": cores, "kind": "port",
               "sample": f"{b_used} molecules per step of the same workload (T={T}, same dims), 4 timed steps after 1 warm-up",
               "ms_per_step": cms,
               "note": "MLX not installable here; oracle/arcvae_oracle.py torch-CPU fp32 restatement of the reference step, "
                       "pinned to the unmodified reference sources by tests/test_ref_pin.py"}

    h2d = hx.numel() * 4 + hc.numel() * 4 + he.numel() * 4
    line = {"metric": "AR-CVAE train molecules/s", "value": B * world / (ms * 1e-3), "unit": "molecules/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16",
            "data": "synthetic", "config": workload_config(args, B, world),
            "e2e": {"value": B * world / (ms_e2e * 1e-3), "unit": "molecules/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
            "gpu_launches": launches, "gpu_launches_per_step": launches // args.steps, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "final_loss": loss_val, **extra}
    if not args.no_sampler:
        try:
            line["sampler"] = bench_sampler(M, vae, cpu=not args.no_cpu)
        except Exception as e:  # noqa: BLE001 — the training line must still be printed
            line["sampler"] = {"error": repr(e)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_fp32(M, B, dx, dc, de, tf_mask, beta, tf, steps=2):
    """The same step with precision='fp32' (every contraction as fp32 FFMA tiles, full-precision activations)."""
    import torch
    vae = M.ARCVAE(**DIMS, seed=67, precision="fp32")
    tr = M.ARCVAETrainerWithLoss(vae.encoder, vae.decoder, None, None, learning_rate=LR, batch_size=B,
                                 lambda_prop=HYPER["lambda_prop"], lambda_collapse=HYPER["lambda_collapse"],
                                 free_bits=HYPER["free_bits"], lambda_mi=HYPER["lambda_mi"])
    tr.train_step(dx, dc, beta, tf, eps=de, tf_mask=tf_mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        d = tr.train_step(dx, dc, beta, tf, eps=de, tf_mask=tf_mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"ms_per_step": ms, "value": B / (ms * 1e-3), "unit": "molecules/s", "steps": steps, "final_loss": float(d["total_loss"])}
    del vae, tr
    torch.cuda.empty_cache()
    return out


def bench_c1(M, precision, n_molecules=4096, batch=64, steps=40):
    """BASELINE configs[0] on the GPU: default model, batch 64, batches drawn from MoleculeDataset.to_batches exactly as
    trainer.py:303 does, data = a synthetic stand-in with the schema of mlx_data/chembl_cns_selfies.json
    (train.py:79-124; the real file is absent from the reference mount).  Host coins (decoder.py:180) are drawn per step."""
    import numpy as np
    import torch
    from mlx_vae_b200.data import load_splits, make_chembl_standin
    np.random.seed(67)                                                        # train.py:75
    train, val, _ = load_splits(make_chembl_standin(n_molecules, max_length=T))
    vae = M.ARCVAE(**DIMS, seed=67, precision=precision)
    tr = M.ARCVAETrainerWithLoss(vae.encoder, vae.decoder, None, train, learning_rate=LR, batch_size=batch,
                                 lambda_prop=HYPER["lambda_prop"], lambda_collapse=HYPER["lambda_collapse"],
                                 free_bits=HYPER["free_bits"], lambda_mi=HYPER["lambda_mi"])
    beta, tfr = tr.compute_beta(0), tr.compute_teacher_forcing_ratio(0, 30)
    it = train.to_batches(batch, shuffle=True)
    for _ in range(5):
        m, c = next(it)
        tr.train_step(m, c, beta, tfr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    n = 0
    for m, c in it:
        if n >= steps or m.shape[0] < batch:
            break
        d = tr.train_step(m, c, beta, tfr)
        n += 1
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / max(1, n)
    tr.check_device_error()
    val_loss = tr._validate(val, beta)
    return {"workload": f"default model, batch {batch} x T={T} through MoleculeDataset.to_batches on a ChEMBL-schema synthetic "
                        f"stand-in of {n_molecules} molecules (80/10/10 split, train-set normalisation), teacher forcing "
                        f"{tfr} with host coins", "ms_per_step": ms, "value": batch / (ms * 1e-3), "unit": "molecules/s",
            "steps": n, "wall_ms_per_step": 1e3 * wall / max(1, n), "final_loss": float(d["total_loss"]),
            "val_loss_tf0": val_loss["loss"]}


if __name__ == "__main__":
    main()
