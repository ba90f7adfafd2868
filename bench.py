#!/usr/bin/env python
"""bench.py -- real-time factor of the Qwen3-ASR hot path on B200 (BASELINE.json metric).

A "step" = one pass of the hot path (transcribe() steps 2-8, /root/reference/src/inference.rs:94-200:
f32 samples -> mel -> encoder -> prefill -> greedy decode -> token ids) over one batch of synthetic
30 s / 16 kHz clips: ONE clip per GPU (BASELINE.json configs[1]; weak scaling over --gpus).
Synthetic weights of the Qwen3-ASR-0.6B architecture never emit EOS, so the decode length is fixed
at --new-tokens (SURVEY.md section 8d "fixed 128 new tokens").

  value : whole-job RTF with the samples already resident in HBM (no H2D in the timed region)
  e2e   : same metric through the public host-buffer API (pinned H2D of the samples and D2H of the
          ids inside the timed region, plus the NCCL gather of ids when N > 1)
  roofline : the decode step (HBM-bound): algorithmic bytes / CUDA-event duration vs measured peak
  cpu_baseline : the oracle (CPU restatement of the reference's tch-CPU path) timed on this box

--impl reference times only the CPU restatement (the Rust reference cannot be built: no cargo).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIP_SECONDS = 30.0


def load_traffic():
    """dram bytes per decode-step launch from the committed `ncu --set full` capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "r01_decode_step_ncu_summary.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return float(d["dram_bytes_read"] + d["dram_bytes_write"])
    except (OSError, KeyError, ValueError):
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def decode_step_bytes(cfg, ctx_tokens: float, batch: int = 1) -> float:
    """Algorithmic HBM bytes of one decoder forward step: every weight once
    (bf16) + the KV cache of `ctx_tokens` tokens (fp32 in parity mode) + the KV append."""
    t = cfg.text
    q_dim, kv_dim = t.num_attention_heads * t.head_dim, t.num_key_value_heads * t.head_dim
    per_layer = (q_dim + 2 * kv_dim) * t.hidden_size + t.hidden_size * q_dim + 3 * t.intermediate_size * t.hidden_size
    weights = (t.num_hidden_layers * per_layer + t.vocab_size * t.hidden_size) * 2
    kv = batch * t.num_hidden_layers * 2 * kv_dim * (ctx_tokens + 1) * 4
    return float(weights + kv)


class ClockSampler:
    def __init__(self, device: int):
        self.device = device
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.device), "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(new_tokens: int, reps: int):
    """The oracle timed on the host cores: one 30 s clip, `new_tokens` greedy tokens, batch 1,
    exactly as the reference would run it (lm_head over all prefill rows included)."""
    import torch
    from oracle import oracle as O
    from qwen3_asr_rs_b200 import synth
    if os.environ.get("OMP_NUM_THREADS") == "1" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # torchrun pins OMP_NUM_THREADS=1 per rank; the CPU arm runs on rank 0 alone and may use the host's cores
        torch.set_num_threads(max(1, min(os.cpu_count() or 1, 64)))
    cfg = O.cfg_0p6b()
    model = O.OracleModel(cfg, synth.make_weights(cfg, 1))
    x = synth.make_clip(0, CLIP_SECONDS)
    times, last = [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        last = O.transcribe_ids(model, x, max_new_tokens=new_tokens)
        times.append(time.perf_counter() - t0)
    return times, last, torch.get_num_threads()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--new-tokens", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--decode", default=None, choices=[None, "mega", "phases"])
    ap.add_argument("--gemm", default=None, choices=[None, "tc", "simt"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = args.steps, max(args.warmup, 3) if args.impl == "ours" else args.warmup
    workload = "Qwen3-ASR-0.6B, batch=1 per GPU, single 30 s 16 kHz clip, greedy decode"
    config = {"workload": workload, "clips_per_gpu": 1, "clip_seconds": CLIP_SECONDS, "new_tokens": args.new_tokens,
              "weights": "synthetic bf16 (seed 1), Qwen3-ASR-0.6B architecture", "activations": "fp32-exact (bf16x3 split / fp32)",
              "l2": "inputs larger than L2: 1.19 GB of weights streamed every decode step",
              "parallelism": f"dp{world}"}

    if args.impl == "reference":
        if rank != 0:
            return
        reps = max(1, K)
        for _ in range(max(0, min(W, 1))):
            cpu_reference_run(min(args.new_tokens, 8), 1)
        times, _, cores = cpu_reference_run(args.new_tokens, reps)
        T = sum(times)
        v = CLIP_SECONDS * reps / T
        line = {"impl": "reference", "metric": "real-time factor (audio-sec/wall-sec)", "value": v, "unit": "x realtime",
                "n_gpus": args.gpus, "steps": reps, "warmup": min(W, 1), "ms_per_step": 1e3 * T / reps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": v, "unit": "x realtime", "cores": cores, "kind": "port",
                                 "sample": f"{reps} x one 30 s clip, {args.new_tokens} new tokens, PyTorch-CPU fp32 restatement "
                                           "of the reference's tch-CPU path (Rust toolchain absent)"},
                "e2e": {"value": v, "unit": "x realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import numpy as np
    import torch
    from qwen3_asr_rs_b200 import AsrInference, config_0p6b, parallel, synth

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = config_0p6b()
    weights = synth.make_weights(cfg, 1)
    eng = AsrInference.from_weights(cfg, weights, device=local_rank)
    del weights
    if args.decode:
        eng.set_option("decode", args.decode)
    if args.gemm:
        eng.set_option("gemm", args.gemm)
    clip = synth.make_clip(rank, CLIP_SECONDS)
    dev = torch.device("cuda", local_rank)

    def one_step():
        r = eng.transcribe_ids([clip], max_new_tokens=args.new_tokens)
        if dist is not None:   # the path's only collective: gather of decoded ids (NCCL over NVLink)
            parallel.gather_token_ids(r.ids, world, args.new_tokens, device=dev)
        return r

    for _ in range(W):
        last = one_step()
    # ---- e2e: host buffers in, ids out ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    stage = {}
    launches = 0
    dec_ms, dec_steps = 0.0, 0
    for _ in range(K):
        last = one_step()
        for k, v in last.stage_ms.items():
            stage[k] = stage.get(k, 0.0) + v / K
        launches = last.kernels_launched
        dec_ms += last.stage_ms["decode"]; dec_steps += last.decode_steps
    barrier()
    t_e2e = time.perf_counter() - t0
    # ---- value: samples resident in HBM ----
    eng.set_option("resident", "1")
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        last = one_step()
    barrier()
    t_val = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    eng.set_option("resident", "0")
    if dist is not None:
        tt = torch.tensor([t_e2e, t_val], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e, t_val = float(tt[0]), float(tt[1])
    ids = last.ids[0]
    if rank == 0:
        audio_total = CLIP_SECONDS * world * K
        value, e2e = audio_total / t_val, audio_total / t_e2e
        peak, peak_kind = load_peaks()
        S = 390 + 15
        ctx_avg = S + (args.new_tokens - 1) / 2.0
        step_bytes = decode_step_bytes(cfg, ctx_avg)
        step_s = (dec_ms / 1e3) / max(dec_steps, 1)
        achieved = step_bytes / step_s / 1e9
        line = {"metric": "real-time factor (audio-sec/wall-sec) Qwen3-ASR-0.6B 30s clips", "value": value, "unit": "x realtime",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * t_val / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (bf16 weights, fp32-exact activations)", "data": "synthetic",
                "config": config,
                "e2e": {"value": e2e, "unit": "x realtime", "h2d_bytes_per_step": int(clip.nbytes),
                        "d2h_bytes_per_step": int(4 * (args.new_tokens + 1)), "ms_per_step": 1e3 * t_e2e / K},
                "gpu_launches": int(launches) * K * 2 + 0,
                "stage_ms": {k: round(v, 4) for k, v in stage.items()},
                "decode": {"steps_per_clip": dec_steps // K, "us_per_step": 1e6 * step_s, "tokens": len(ids)},
                "roofline": {"kernel": "decoder forward step (batch 1)", "bound": "hbm", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "peak_kind": peak_kind,
                             "bytes_per_launch": step_bytes, "traffic": load_traffic(),
                             "traffic_source": "profiles/r01_decode_step_ncu_summary.json (ncu --set full, 1 launch)"},
                "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            times, ref, cores = cpu_reference_run(args.new_tokens, 1)
            line["cpu_baseline"] = {"value": CLIP_SECONDS / times[0], "unit": "x realtime", "cores": cores, "kind": "port",
                                    "sample": f"one 30 s clip, {args.new_tokens} new tokens (the whole step), PyTorch-CPU fp32 "
                                              "restatement of the reference's tch-CPU path",
                                    "ids_match_gpu": ref.ids == ids}
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
