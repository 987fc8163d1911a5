#!/usr/bin/env python
"""bench.py -- throughput of the scan path on B200 (text GB/s scanned + matches/s), next to the reference's CPU loop.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3|c2|c4s|c5] [--impl ours|reference]

A "step" is one pass of the hot path (acm_b200_scan_ex: every kernel of the engine) over this rank's shard of the
synthetic text.  `value` is measured with the shard resident in HBM; `e2e` goes through the same C-ABI call with host
(pinned) buffers, H2D of the text and D2H of the records inside the timed region.  One process per GPU; shards are
independent (weak scaling: every rank scans --gib GiB), the only collective is the all-gather of the match counts.
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

CONFIGS = {
    # name: (description, dictionary builder args, text kind, default GiB per GPU)
    "c3": dict(workload="BASELINE configs[2]: 100k random-byte patterns (len 4-32) over 8 GiB synthetic bytes, keywords planted every 4 KiB", nb_patterns=100_000, kind=0, gib=8.0),
    "c4s": dict(workload="BASELINE configs[3] single-GPU slice: 1M random-byte patterns (len 4-32) over 2 GiB synthetic bytes per GPU", nb_patterns=1_000_000, kind=0, gib=2.0),
    "c5": dict(workload="BASELINE configs[4] slice: uint32 token ids, Zipf(1.0) over a 50k vocabulary, 200k n-gram keywords (len 2-8) cut from the stream, 64 Mi tokens per GPU",
               nb_patterns=200_000, kind=5, gib=0.25, width=4),
    "c2": dict(workload="BASELINE configs[1]: 1k most frequent English words of the novel over 1 GiB synthetic printable ASCII, keywords planted every 4 KiB", nb_patterns=0, kind=1, gib=1.0),
}
TEXT_SEED, DICT_SEED, PLANT_SEED, PLANT_PERIOD = 0xC0FFEE, 0xD1C7, 0x5EED, 4096


def zipf_tokens(n, first, vocab=50_000, seed=0xC0FFEE):
    """Zipf(1.0) token ids by inverse CDF; block-seeded so that any rank can generate its own shard."""
    cdf = np.cumsum(1.0 / np.arange(1, vocab + 1))
    cdf /= cdf[-1]
    out = np.empty(n, dtype=np.uint32)
    blk = 1 << 22
    for b0 in range(first // blk, (first + n + blk - 1) // blk):
        r = np.random.default_rng([seed, b0]).random(blk)
        toks = np.searchsorted(cdf, r).astype(np.uint32)
        lo, hi = max(first, b0 * blk), min(first + n, (b0 + 1) * blk)
        out[lo - first:hi - first] = toks[lo - b0 * blk:hi - b0 * blk]
    return out


def build_dictionary(cfg):
    if cfg.get("width", 1) == 4:
        stream = zipf_tokens(1 << 22, 0)
        rng = np.random.default_rng(DICT_SEED)
        starts, lens = rng.integers(0, len(stream) - 8, size=cfg["nb_patterns"]), rng.integers(2, 9, size=cfg["nb_patterns"])
        flat = np.concatenate([stream[a:a + l] for a, l in zip(starts, lens)])
        return flat, np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    if cfg["nb_patterns"]:
        rng = np.random.default_rng(DICT_SEED)
        lens = rng.integers(4, 33, size=cfg["nb_patterns"])
        flat = rng.integers(0, 256, size=int(lens.sum())).astype(np.uint8)
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        return flat, offsets
    import gzip
    import re
    from collections import Counter

    novel = gzip.open(os.path.join(ROOT, "tests", "golden", "mrs_dalloway.txt.gz"), "rb").read()
    cnt = Counter(re.findall(rb"[a-z]+", novel.lower()))
    words = [w for w, _ in sorted(cnt.items(), key=lambda kv: (-kv[1], kv[0]))[:1000]]
    flat = np.frombuffer(b"".join(words), dtype=np.uint8)
    offsets = np.concatenate([[0], np.cumsum([len(w) for w in words])]).astype(np.uint64)
    return flat, offsets


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md 'clocks' line)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        """Samples that arrived inside [t_begin, t_end] (the timed region); if fewer than 3, every sample taken under load
        (warm-up + timed region), flagged in `window`."""
        if self.proc:
            self.proc.terminate()
        inside = [r for t, r in self.rows if t_begin is not None and t_begin <= t <= t_end + 0.05]
        window = "timed region"
        rows = inside
        if len(rows) < 3:
            rows, window = [r for _, r in self.rows], "warm-up + timed region (timed region shorter than the sampler's period)"
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        self.window = window
        for r in rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm), "window": window}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_oracle(flat, offsets, width=1):
    """The reference's own CPU implementation of the path (oracle/_ref, classic build: its Meyer build needs minutes to ingest 100k
    patterns), else the C restatement."""
    from oracle import pyoracle

    for kind, label in (("ref_classic", "reference"), ("ref_meyer", "reference"), ("port", "port")):
        if pyoracle.available(kind):
            o = pyoracle.Oracle(kind, width)
            o.insert_many(flat=flat, offsets=offsets)
            return o, label, kind
    raise RuntimeError("no oracle library is built")


def host_text(ac75, cfg, nb, first, flat, offsets):
    if cfg.get("width", 1) == 4:
        return zipf_tokens(nb, first)
    return ac75.generate_text(nb, first=first, kind=cfg["kind"], seed=TEXT_SEED, plant_seed=PLANT_SEED, plant_period=PLANT_PERIOD, dict_flat=flat, dict_offsets=offsets)


def run_reference(args, cfg, rank, world):
    """--impl reference: the reference's CPU loop (acm_match + acm_get_match per symbol) on all host cores, bounded sample per step."""
    if rank != 0:
        return
    ac75 = entry.load_package()
    flat, offsets = build_dictionary(cfg)
    width = cfg.get("width", 1)
    oracle, label, kind = reference_oracle(flat, offsets, width)
    cores = host_cores()
    sample = int(min(cores, 64) * (1 << 20) * (4 if cfg["nb_patterns"] == 0 else 1)) // width
    text = host_text(ac75, cfg, sample, 0, flat, offsets)
    threads = min(cores, 64)
    for _ in range(max(args.warmup, 1)):  # also triggers the classic build's lazy fail-link construction
        oracle.scan_mt(text[: max(sample // 8, 1 << 16)], threads)
    secs, matches = [], 0
    for _ in range(args.steps):
        m, s = oracle.scan_mt(text, threads)
        secs.append(s)
        matches = m
    t = float(np.mean(secs))
    gbs = sample * width / t / 1e9
    line = {"impl": "reference", "metric": "text_GB_per_s_scanned", "value": gbs, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8" if width == 1 else "u32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "sample": f"first {sample * width >> 20} MiB of the same generated text per step"},
            "matches_per_s": matches / t,
            "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": label, "sample": f"{kind}: acm_match+acm_get_match loop over the first {sample * width >> 20} MiB, {threads} threads (one cursor each)"},
            "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gib", type=float, default=None, help="GiB of text per GPU (default: the config's size)")
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--option", action="append", default=[], help="key=value passed to acm_b200_set_option")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    cfg = dict(CONFIGS[args.config])
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist

    ac75 = entry.load_package()
    ac75.lib()  # fails loudly if libac75.so is missing
    if not torch.cuda.is_available() or ac75.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the scan path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    gib = args.gib if args.gib is not None else cfg["gib"]
    width = cfg.get("width", 1)
    shard = int(gib * (1 << 30)) // width // 4096 * 4096  # symbols per GPU
    flat, offsets = build_dictionary(cfg)
    t0 = time.time()
    m = ac75.Machine(width)
    m.insert_many(flat=flat, offsets=offsets)
    build_s = time.time() - t0
    if args.engine != "auto":
        m.set_option("engine", args.engine)
    for kv in args.option:
        k, v = kv.split("=", 1)
        m.set_option(k, v)
    m.finalise(local)
    lmax = m.max_keyword_length
    lead = 0 if rank == 0 else -(-(lmax - 1) // 16) * 16
    first = rank * shard - lead  # weak scaling: rank g owns [g*shard, (g+1)*shard) of a world*shard text
    n = shard + lead

    stream = torch.cuda.Stream()  # the library launches every kernel of the scan on this stream; the events below are recorded on it
    torch.cuda.set_stream(stream)
    d_text = torch.empty(n * width + 64, dtype=torch.uint8, device="cuda")
    if width == 4:
        d_text[: n * 4].copy_(torch.from_numpy(zipf_tokens(n, first).view(np.uint8)))
    else:
        ac75.generate_text(n, first=first, kind=cfg["kind"], seed=TEXT_SEED, plant_seed=PLANT_SEED, plant_period=PLANT_PERIOD, dict_flat=flat, dict_offsets=offsets,
                           device_ptr=d_text.data_ptr(), stream=stream.cuda_stream)
    # room for every record: sparse configs report ~1 match per 4 KiB; c2 (single-letter words) ~1 per 15 bytes; c5 up to ~1 per token
    cap = {"c2": n // 8, "c5": 2 * n}.get(args.config, max(1 << 20, n // 512))
    d_matches = torch.empty(cap * 16, dtype=torch.uint8, device="cuda")

    def step_device():
        return m.scan_device(d_text.data_ptr(), n, lead=lead, base=first, d_matches_ptr=d_matches.data_ptr(), capacity=cap, stream=stream.cuda_stream)

    counts_in = torch.zeros(1, dtype=torch.int64, pin_memory=True)
    counts_all = torch.zeros(world, dtype=torch.int64, device="cuda")

    def exchange(local_count):
        if world == 1:
            return local_count
        counts_in[0] = local_count
        dist.all_gather_into_tensor(counts_all, counts_in.cuda(non_blocking=True))  # the path's only collective: one match count per GPU
        return int(counts_all.sum().item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)  # nvidia-smi needs a moment before its first sample
    t_load = time.time()
    while True:  # warm-up: at least W steps, and long enough for the clock sampler to see the GPU under load
        for _ in range(max(args.warmup, 1)):
            local_matches = step_device()
            exchange(local_matches)
        if args.warmup == 0 or time.time() - t_load > 0.25:
            break
    assert args.warmup == 0 or local_matches <= cap, f"record buffer too small: {local_matches} > {cap}"
    st0 = m.stats()
    barrier()
    t_begin = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main_ms, kernel_ms, cands = [], [], 0
    ev0.record(stream)
    for _ in range(args.steps):
        local_matches = step_device()
        total_matches = exchange(local_matches)
        s = m.stats()
        main_ms.append(s["main_kernel_ms"])
        kernel_ms.append(s["scan_kernel_ms"])
        cands = s["last_nb_candidates"]
    ev1.record(stream)
    barrier()
    clocks = sampler.stop(t_begin, time.time())
    st1 = m.stats()
    print(f"[rank {rank}] ms/step(local)={ev0.elapsed_time(ev1) / max(args.steps, 1):.3f} kernel_ms={np.mean(kernel_ms):.3f} main_ms={np.mean(main_ms):.3f} "
          f"fallbacks={st1['fallback_count']} stride={st1.get('filter_stride')} matches={local_matches} cands={cands}", file=sys.stderr)
    ms_step = ev0.elapsed_time(ev1) / max(args.steps, 1)
    if world > 1:
        t = torch.tensor([ms_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item())
    total_bytes = shard * width * world  # bytes of the symbols owned (the lead overlap is overhead, not counted)
    value = total_bytes / (ms_step * 1e-3) / 1e9

    # end to end: host (pinned) text in, host (pinned) records out, through the same C-ABI call
    e2e = None
    if not args.no_e2e:
        # with several ranks on one host the pinned buffers are capped at 2 GiB per rank (8 x 8 GiB of pinned memory is not needed to
        # measure a PCIe-bound rate); the value stays bytes scanned / time, over what each rank really copies and scans
        n_e2e = n if world == 1 else min(n, (2 << 30) // width // 4096 * 4096)
        h_text = torch.empty(n_e2e * width, dtype=torch.uint8, pin_memory=True)
        h_text.copy_(d_text[: n_e2e * width])
        h_out = torch.empty(cap * 16, dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()
        e2e_steps = max(2, min(args.steps, 5))
        m.scan_host_to_host(h_text.data_ptr(), n_e2e, h_out.data_ptr(), cap, lead=lead, base=first)  # warm-up (allocates the staging buffer)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            got = m.scan_host_to_host(h_text.data_ptr(), n_e2e, h_out.data_ptr(), cap, lead=lead, base=first)
            exchange(got)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        if world > 1:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        assert got <= cap and (n_e2e != n or got == local_matches), (got, local_matches, cap)
        e2e = {"value": (n_e2e - lead) * width * world / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(n_e2e * width),
               "d2h_bytes_per_step": int(min(got, cap) * 16 + 8), "ms_per_step": e2e_ms, "steps": e2e_steps, "bytes_scanned_per_rank": int((n_e2e - lead) * width)}
        del h_text, h_out

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    t_main = float(np.mean(main_ms)) * 1e-3
    algo_bytes = n * width + local_matches * 16  # N*w + M*16 per launch of this rank
    achieved = algo_bytes / t_main / 1e9
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    line = {
        "metric": "text_GB_per_s_scanned", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8" if width == 1 else "u32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "bytes_per_gpu": shard * width, "engine": st1["engine"], "nb_keywords": st1["nb_keywords"], "nb_states": st1["nb_states"],
                   "l2": "inputs larger than L2 (no flush needed)", "text_seed": hex(TEXT_SEED), "dict_seed": hex(DICT_SEED), "table_bytes": st1["table_bytes"],
                   "smem_bytes": st1["smem_bytes"], "dictionary_build_s": round(build_s, 2), "finalise_ms": round(st1["finalise_ms"], 1),
                   "filter_stride": st1.get("filter_stride"), "filter_hit_rate": round(st1["filter_fp"], 5), "fallback_count": st1["fallback_count"]},
        "matches_per_s": total_matches / (ms_step * 1e-3), "matches_per_step": total_matches, "candidates_per_step_rank0": cands,
        "kernel_ms": {"main": float(np.mean(main_ms)), "all": float(np.mean(kernel_ms))},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": prof.get(args.config), "peak_source": peak_src,
                     "kernel": ("filter_scan_s2_kernel" if st1.get("filter_stride") == 2 else "filter_scan_kernel") if st1["engine"] == "filter" else "dfa_scan_kernel (count + event recording) + dfa_emit_events_kernel" if st1.get("dfa_event_scans") else "dfa_scan_kernel (count) + dfa_emit_kernel (walking emit)",
                     "algorithmic_bytes_per_launch": int(algo_bytes)},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(st1["total_kernel_launches"] - st0["total_kernel_launches"]),
    }
    if not args.no_cpu_baseline and world == 1:
        oracle, label, kind = reference_oracle(flat, offsets, width)
        sample = ((2 << 20) if cfg["nb_patterns"] else (32 << 20)) // width
        text = host_text(ac75, cfg, sample, 0, flat, offsets)
        oracle.count(text[: 1 << 16])  # classic build: lazy construction of the fail links outside the timed region
        oracle.reset_cursor()
        t0 = time.perf_counter()
        cm = oracle.count(text)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": sample * width / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": label,
                                "sample": f"{kind}: acm_match+acm_get_match loop over the first {sample * width >> 20} MiB of the same text, {cm} matches, {dt:.1f} s; host has {host_cores()} cores"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
