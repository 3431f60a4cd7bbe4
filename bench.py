#!/usr/bin/env python
"""Benchmark of the hot path: effective decoded tokens/sec (post-verify), Whisper-large-v2 + 10
Medusa-Linear heads, batch 1, 30 s synthetic audio  (BASELINE.json metric; configs[1]).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference            # the CPU oracle (port of the reference) on host cores

One "step" = one 30 s clip through the whole path (PCM -> log-mel -> encoder -> cross-K/V ->
Medusa speculative loop to max_length).  Streams are independent (reference asserts batch 1,
model.py:1451), so N GPUs run N streams per step with no data-path collective ("weak" scaling);
the only collective is the NCCL broadcast of the packed weights at start-up.

Printed keys (one JSON line, rank 0):
  value       tokens/s of the decode loop with the encoder output resident in HBM (device time,
              CUDA events on the engine stream) -- whole job: sum of tokens over ranks / max time
  e2e         same metric through WhisperMedusaModel.generate_from_pcm with HOST buffers: pinned
              PCM in, token ids out, H2D/D2H inside the timed region, wall clock max over ranks
  roofline    decode iteration vs the HBM roofline: algorithmic bytes of SURVEY.md 8(d)
  cpu_baseline the CPU oracle timed on this box's host cores on a bounded sample
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
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "effective decoded tokens/sec (post-verify) Whisper-large-v2+Medusa"


def ncu_traffic_bytes(args):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/r1b_ring_kernel_ncu.json) -- only meaningful for the configuration that
    capture was taken on (large-v2, 10 linear heads, persistent mode); null otherwise."""
    if not (args.preset == "large-v2" and args.heads == 10 and args.heads_type == "base_head" and args.mode == "persistent"):
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "r1b_ring_kernel_ncu.json")) as f:
            pl = json.load(f)["per_launch"]
        return int(pl["dram_bytes_read"]) + int(pl["dram_bytes_write"])
    except Exception:
        return None


def algorithmic_bytes(cfg, iterations: int, sweeps_a: int, n_mean: float):
    """fp16 bytes the decode loop must move, from the unit figures of SURVEY.md 8(d)
    (DESIGN.md section 4).  Returns (bytes of THIS engine's schedule, bytes of the reference schedule).

    reference schedule: two decoder sweeps per iteration (pass A + pass B), K+1 head matrices + the
    base head again, two vocabulary projections.  Engine schedule (sweep elision): one sweep per
    iteration plus one extra sweep per accept-0 iteration and for the prompt (`sweeps_a`)."""
    d, f, V, K, S = cfg.d_model, cfg.decoder_ffn_dim, cfg.vocab_size, cfg.medusa_num_heads, cfg.max_source_positions
    layer = (6 * d * d + 2 * d * f) * 2 + 2 * S * d * 2          # weights + cross K/V of one layer
    proj = V * d * 2
    head = (d * d + d) * 2
    n_l = cfg.decoder_layers + (1 if cfg.is_block else 0)
    kv_sweep = 2 * d * 2 * n_l * n_mean                            # self-K/V rows read by one sweep
    sweep = n_l * layer + kv_sweep
    heads_iter = (K * head) if cfg.is_block else ((K + 1) * head + head)
    fixed = 2 * proj + heads_iter
    ref = iterations * (2 * sweep + fixed)
    eng = iterations * (sweep + fixed) + sweeps_a * sweep
    if cfg.is_block:                                               # tail re-runs the block layer on one row
        eng += iterations * (layer + kv_sweep / n_l)
    return float(eng), float(ref)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def pick_threads(cfg) -> int:
    """PyTorch eager on every hardware thread of a large host is far slower than on a subset (the
    batch-1 decode is DRAM-bound and the thread barrier cost grows); time one vocabulary projection
    for a few thread counts and keep the fastest.  `cores` in the JSON is what was actually used."""
    n = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, n) if c <= n})
    w = torch.randn(cfg.vocab_size, cfg.d_model)
    x = torch.randn(1, cfg.d_model)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        for _ in range(2):
            torch.nn.functional.linear(x, w)
        t0 = time.perf_counter()
        for _ in range(5):
            torch.nn.functional.linear(x, w)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    return best


def cpu_reference_sample(cfg, sd, pcm, max_iters: int, threads: int):
    """Time the CPU oracle (PyTorch eager fp32 restatement of the reference loop: one proj_out per
    head, two passes per iteration, KV concatenation) on a bounded sample of the workload."""
    from oracle import medusa_ref as M
    from oracle import whisper_ref as W

    torch.set_num_threads(threads)
    w = W.RefWeights(sd)
    t0 = time.perf_counter()
    mel = torch.from_numpy(W.log_mel_spectrogram(pcm))
    t1 = time.perf_counter()
    with torch.inference_mode():
        enc = W.encoder_forward(w, cfg, mel, "fp32")
        t2 = time.perf_counter()
        prompt = M.init_tokens(cfg, "en" if cfg.is_multilingual else None)
        gp = M.gen_params(cfg, prompt)
        tr = M.medusa_greedy_search(w, cfg, enc, prompt, gp, "fp32", max_iters=max_iters)
    t3 = time.perf_counter()
    n_tok = len(tr.sequences) - len(prompt)
    return {"tokens": n_tok, "iters": tr.iters, "s_mel": t1 - t0, "s_encoder": t2 - t1, "s_decode": t3 - t2,
            "tok_s_decode": n_tok / (t3 - t2), "tok_s_e2e": n_tok / (t3 - t0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--mode", default=os.environ.get("WM_DECODE_MODE", "persistent"), choices=["graph", "persistent_simple", "persistent"])
    ap.add_argument("--preset", default="large-v2")
    ap.add_argument("--heads", type=int, default=10)
    ap.add_argument("--heads-type", default="base_head", choices=["base_head", "medusa_block"])
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--cpu-iters", type=int, default=8, help="speculative iterations of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    from whisper_medusa_b200.synthetic import preset_config, synthetic_audio, synthetic_state_dict

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = preset_config(args.preset, heads=args.heads, heads_type=args.heads_type)
    workload = (f"whisper-{args.preset} + {args.heads} Medusa-{'Block' if cfg.is_block else 'Linear'} heads, batch 1, "
                f"{args.seconds:g} s synthetic 16 kHz audio per stream, greedy/typical acceptance, max_length 448")
    config = {"workload": workload, "streams_per_step": world, "parallelism": f"replicas x{world} (independent streams)",
              "weights": "seeded synthetic fp16 (seed 0)", "l2": "weights (3.1 GB) >> L2: every iteration re-streams them from HBM"}
    threads = pick_threads(cfg) if (args.impl == "reference" or not args.no_cpu_baseline) else (os.cpu_count() or 1)

    # ------------------------------------------------------------------ reference arm (CPU oracle)
    if args.impl == "reference":
        if rank != 0:
            return
        sd = synthetic_state_dict(cfg, seed=0)
        vals, per_step = [], []
        for i in range(args.warmup + args.steps):
            pcm = synthetic_audio(args.seconds, stream_id=i)
            t0 = time.perf_counter()
            r = cpu_reference_sample(cfg, sd, pcm, args.cpu_iters, threads)
            if i >= args.warmup:
                per_step.append(time.perf_counter() - t0)
                vals.append(r)
        tok = sum(v["tokens"] for v in vals)
        dec = sum(v["s_decode"] for v in vals)
        e2e = tok / sum(per_step)
        sample = (f"per step: 1 clip, log-mel + full encoder + first {args.cpu_iters} speculative iterations "
                  f"({vals[0]['tokens']} tokens) of the same workload")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": tok / dec, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(per_step) / len(per_step),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": tok / dec, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ------------------------------------------------------------------ engine arm
    from whisper_medusa_b200 import WhisperMedusaModel

    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = torch.device("cuda", local_rank)
    sd = synthetic_state_dict(cfg, seed=0) if (rank == 0 or world == 1) else None
    model = WhisperMedusaModel(cfg, sd)
    model.to(device, broadcast_src=0 if world > 1 else None)   # NCCL broadcast of the packed blob (N > 1)
    model.set_decode_mode(args.mode)

    def clip(i):
        return torch.from_numpy(synthetic_audio(args.seconds, stream_id=rank + world * i)).pin_memory()

    clips = [clip(i) for i in range(args.warmup + args.steps)]
    language = "en" if cfg.is_multilingual else None

    def barrier():
        torch.cuda.synchronize(device)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)

    for i in range(args.warmup):
        model.generate_from_pcm(clips[i], language=language)
    barrier()
    toks = iters = launches = 0
    dec_ms = enc_ms = mel_ms = 0.0
    n_sum = 0.0
    sweeps_a = 0
    with ClockSampler(local_rank) as cs:
        t0 = time.perf_counter()
        for i in range(args.warmup, args.warmup + args.steps):
            out = model.generate_from_pcm(clips[i], language=language)      # H2D of PCM + D2H of ids inside
            _ = out.cpu()
            tr = model.last_trace
            toks += tr.n_new_tokens
            iters += tr.iterations
            dec_ms += tr.ms_decode
            enc_ms += tr.ms_encoder
            mel_ms += tr.ms_mel
            launches += tr.launches_encode + tr.launches_decode
            n_sum += 0.5 * len(tr.sequences) * tr.iterations   # mean self-KV length over the run ~ L_final / 2
            sweeps_a += 1 + sum(1 for a in tr.accept_lengths[:-1] if a == 0)   # prompt + one per accept-0 iteration
        barrier()
        wall = time.perf_counter() - t0
    clocks = cs.summary()

    stats = torch.tensor([toks, iters, dec_ms, wall, enc_ms, mel_ms, launches, n_sum, sweeps_a], dtype=torch.float64, device=device)
    if dist is not None:
        allv = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(allv, stats)
        allv = torch.stack(allv).cpu()
    else:
        allv = stats.cpu()[None]
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    tot_tok = float(allv[:, 0].sum())
    tot_iter = float(allv[:, 1].sum())
    max_dec_ms = float(allv[:, 2].max())
    max_wall = float(allv[:, 3].max())
    value = tot_tok / (max_dec_ms / 1e3)
    e2e_value = tot_tok / max_wall
    # roofline of the decode iteration (the dominant kernel: one launch per iteration in persistent mode)
    it0, n_mean = float(allv[0, 1]), float(allv[0, 7]) / max(1.0, float(allv[0, 1]))
    bytes_eng, bytes_ref = algorithmic_bytes(cfg, int(it0), int(allv[0, 8]), n_mean)
    if args.mode == "graph":
        pass  # same schedule; the graph mode only differs in how stages are launched
    ms_iter = float(allv[0, 2]) / max(1.0, it0)
    achieved = bytes_eng / (float(allv[0, 2]) / 1e3) / 1e9
    peak, peak_src = measured_peak_gbs()
    tok_iter = tot_tok / max(1.0, tot_iter)
    ref_roofline_tok_s = peak * 1e9 / (bytes_ref / max(1.0, it0)) * tok_iter   # per GPU
    result = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * max_wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16 weights+KV / fp32 accumulate", "data": "synthetic", "config": dict(config, decode_mode=args.mode),
        "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": 480000 * 4,
                "d2h_bytes_per_step": int(4 * tot_tok / max(1, world * args.steps)) + 16 * 4},
        "gpu_launches": int(allv[:, 6].sum()),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic_bytes(args), "peak_source": peak_src,
                     "kernel": "dec_iteration_ring_kernel (one launch = one speculative iteration)"
                     if args.mode == "persistent" else f"decode iteration ({args.mode})",
                     "algorithmic_bytes_per_iteration": bytes_eng / max(1.0, it0), "ms_per_iteration": ms_iter,
                     "reference_schedule_bytes_per_iteration": bytes_ref / max(1.0, it0),
                     "frac_of_reference_schedule_roofline": (value / world) / ref_roofline_tok_s},
        "detail": {"tokens_per_step": tot_tok / (world * args.steps), "iterations_per_step": tot_iter / (world * args.steps),
                   "tokens_per_iteration": tot_tok / max(1.0, tot_iter), "ms_mel": float(allv[0, 5]) / args.steps,
                   "ms_encoder": float(allv[0, 4]) / args.steps, "ms_decode": float(allv[0, 2]) / args.steps},
    }
    if not args.no_cpu_baseline and world >= 1:
        try:
            if sd is None:
                sd = synthetic_state_dict(cfg, seed=0)
            r = cpu_reference_sample(cfg, sd, clips[0].numpy(), args.cpu_iters, threads)
            result["cpu_baseline"] = {
                "value": r["tok_s_decode"], "unit": "tokens/s", "cores": threads, "kind": "port",
                "sample": (f"1 clip: log-mel {r['s_mel']:.2f}s + encoder {r['s_encoder']:.2f}s + first {r['iters']} "
                           f"speculative iterations ({r['tokens']} tokens) in {r['s_decode']:.2f}s; e2e {r['tok_s_e2e']:.1f} tok/s")}
        except Exception as e:  # noqa: BLE001
            result["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(result))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
