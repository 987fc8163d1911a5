#!/usr/bin/env python
"""bench.py -- throughput of the scan path on B200 (text GB/s scanned + matches/s), next to the reference's CPU loop.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3|c2|c4|c4s|c5] [--scaling strong|weak] [--impl ours|reference]

A "step" is one pass of the hot path (acm_b200_scan_ex: every kernel of the engine) over this rank's shard of the
synthetic text.  `value` is measured with the shard resident in HBM; `e2e` goes through the same C-ABI call with host
(pinned) buffers, H2D of the text and D2H of the records inside the timed region.  One process per GPU; shards are
independent, the only collective is the all-gather of the match counts (issued on the scan's stream, read once after the
timed region).  Default workload = BASELINE configs[2] as stated: ONE 8 GiB text split over the N GPUs (`scaling: strong`);
`--scaling weak` gives every rank its own 8 GiB and is also reported next to the strong figure (`weak` object).  At N=1 the
other configs are measured too, at a reduced number of steps, and reported in `extra_configs`.  Rank 0 prints ONE JSON line.
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
sys.path.insert(0, os.path.join(ROOT, "tests", "support"))
import __graft_entry__ as entry  # noqa: E402
import textgen  # noqa: E402  (bench/test support: synthetic text; the host path is numpy, the device path its own small CUDA library)

CONFIGS = {
    "c3": dict(workload="BASELINE configs[2]: 100k random-byte patterns (len 4-32) over 8 GiB synthetic bytes, keywords planted every 4 KiB", nb_patterns=100_000, kind=0, gib=8.0,
               scaling="strong", cpu_sample=64 << 20, cpu_kind="reference"),
    "c4": dict(workload="BASELINE configs[3]: 1M random-byte patterns (len 4-32) over 16 GiB synthetic bytes sharded across 8 GPUs = 2 GiB per GPU", nb_patterns=1_000_000, kind=0,
               gib=2.0, scaling="weak", cpu_sample=8 << 20, cpu_kind="port"),
    "c4s": dict(workload="BASELINE configs[3] single-GPU slice: 1M random-byte patterns (len 4-32) over 2 GiB synthetic bytes per GPU", nb_patterns=1_000_000, kind=0, gib=2.0,
                scaling="weak", cpu_sample=8 << 20, cpu_kind="port"),
    "c5": dict(workload="BASELINE configs[4] slice: uint32 token ids, Zipf(1.0) over a 50k vocabulary, 200k n-gram keywords (len 2-8) cut from the stream, 64 Mi tokens per GPU, "
                        "then 10 Meyer rounds of +2,000 n-grams, each followed by a scan of the next tenth of the stream with the cursor carried",
               nb_patterns=200_000, kind=5, gib=0.25, width=4, scaling="weak", cpu_sample=8 << 20, cpu_kind="port", meyer_rounds=10, meyer_add=2000),
    "c2": dict(workload="BASELINE configs[1]: 1k most frequent English words of the novel over 1 GiB synthetic printable ASCII, keywords planted every 4 KiB", nb_patterns=0, kind=1,
               gib=1.0, scaling="weak", cpu_sample=32 << 20, cpu_kind="reference"),
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


def ngram_dictionary(nb, seed):
    stream = zipf_tokens(1 << 22, 0)
    rng = np.random.default_rng(seed)
    starts, lens = rng.integers(0, len(stream) - 8, size=nb), rng.integers(2, 9, size=nb)
    flat = np.concatenate([stream[a:a + l] for a, l in zip(starts, lens)])
    return flat, np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)


def build_dictionary(cfg):
    if cfg.get("width", 1) == 4:
        return ngram_dictionary(cfg["nb_patterns"], DICT_SEED)
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


def bind_to_gpu_numa_node(local):
    """Pins this rank to the CPUs of its GPU's NUMA node BEFORE any pinned allocation (first touch then places the staging buffers
    next to the GPU's PCIe root).  Returns a short description for the JSON line."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)], capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bus = bus[-12:] if len(bus) > 12 else bus  # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return "numa node unknown (single node)"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"numa node {node}, {len(allowed)} cpus"
        return f"numa node {node}: none of its cpus is in this process's affinity mask"
    except Exception as e:  # no sysfs / no permission: leave the placement to the OS
        return f"not bound ({type(e).__name__})"


def reference_oracle(flat, offsets, width=1, prefer="reference"):
    """The reference's own CPU implementation of the path (oracle/_ref, classic build: its Meyer build needs minutes to ingest 100k
    patterns), else -- or for the dictionaries it cannot ingest in bench time -- the C restatement."""
    from oracle import pyoracle

    order = (("ref_classic", "reference"), ("ref_meyer", "reference"), ("port", "port")) if prefer == "reference" else (("port", "port"),)
    for kind, label in order:
        if pyoracle.available(kind):
            o = pyoracle.Oracle(kind, width)
            o.insert_many(flat=flat, offsets=offsets)
            return o, label, kind
    raise RuntimeError("no oracle library is built")


def host_text(cfg, nb, first, flat, offsets):
    """Host copy of text[first, first + nb): numpy only -- no CUDA, no product library."""
    if cfg.get("width", 1) == 4:
        return zipf_tokens(nb, first)
    return textgen.host_text(nb, first, cfg["kind"], TEXT_SEED, PLANT_SEED, PLANT_PERIOD, flat, offsets)


def run_reference(args, cfg, rank, world):
    """--impl reference: the reference's CPU loop (acm_match + acm_get_match per symbol) on all host cores, bounded sample per step."""
    if rank != 0:
        return
    flat, offsets = build_dictionary(cfg)
    width = cfg.get("width", 1)
    oracle, label, kind = reference_oracle(flat, offsets, width, cfg.get("cpu_kind", "reference"))
    cores = host_cores()
    threads = min(cores, 64)
    per_thread = max((64 << 20) // threads, 1 << 20)  # >= 64 MiB per step over all threads
    sample = threads * per_thread // width
    text = host_text(cfg, sample, 0, flat, offsets)
    for _ in range(max(args.warmup, 1)):  # also triggers the classic build's lazy fail-link construction
        oracle.scan_mt(text[: max(sample // 16, 1 << 16)], threads)
    secs, matches = [], 0
    for _ in range(args.steps):
        m, s = oracle.scan_mt(text, threads)
        secs.append(s)
        matches = m
    t = float(np.mean(secs))
    gbs = sample * width / t / 1e9
    line = {"impl": "reference", "metric": "text_GB_per_s_scanned", "value": gbs, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "u8" if width == 1 else "u32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "sample": f"first {sample * width >> 20} MiB of the same generated text per step"},
            "matches_per_s": matches / t,
            "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": label, "sample": f"{kind}: acm_match+acm_get_match loop over the first {sample * width >> 20} MiB, {threads} threads (one cursor each)"},
            "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def record_hash(torch, d_matches, count):
    """Order-independent 64-bit checksum of `count` 16-byte records in device memory (sum of mixed words, wrapping)."""
    if count == 0:
        return 0
    w = d_matches[: count * 16].view(torch.int64).view(-1, 2)
    h = (w[:, 0] * -7046029254386353131) ^ (w[:, 1] * -4417276706812531889)  # 0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F as int64
    h = h ^ (h >> 29)
    return int(h.sum().item()) & 0xFFFFFFFFFFFFFFFF


def measure(name, cfg, args, ac75, torch, dist, rank, world, local, scaling, steps, warmup, with_e2e, with_cpu, with_check, numa):
    """One config on this rank's GPU; rank 0 gets the JSON object, the others None."""
    width = cfg.get("width", 1)
    gib = args.gib if (args.gib is not None and name == args.config) else cfg["gib"]
    total_syms = int(gib * (1 << 30)) // width // 4096 * 4096
    flat, offsets = build_dictionary(cfg)
    t0 = time.time()
    m = ac75.Machine(width)
    m.insert_many(flat=flat, offsets=offsets)
    build_s = time.time() - t0
    if args.engine != "auto" and name == args.config:
        m.set_option("engine", args.engine)
    if name == args.config:
        for kv in args.option:
            k, v = kv.split("=", 1)
            m.set_option(k, v)
    m.finalise(local)
    lmax = m.max_keyword_length
    if scaling == "strong":  # ONE text of total_syms symbols, split over the ranks (shard.py)
        plan = ac75.plan_shards(total_syms, world, lmax)
        a, b, lead = plan[rank]
        text_syms = total_syms
    else:  # weak: every rank owns total_syms symbols of a world x total_syms text
        lead = 0 if rank == 0 else -(-(lmax - 1) // 16) * 16
        a, b = rank * total_syms, (rank + 1) * total_syms
        text_syms = total_syms * world
        plan = [(g * total_syms, (g + 1) * total_syms, 0 if g == 0 else -(-(lmax - 1) // 16) * 16) for g in range(world)]
    first, n, owned = a - lead, b - a + lead, b - a

    stream = torch.cuda.Stream()  # the library launches every kernel of the scan on this stream; the events below are recorded on it
    torch.cuda.set_stream(stream)

    def device_text(first_sym, nsym):
        d = torch.empty(nsym * width + 64, dtype=torch.uint8, device="cuda")
        if width == 4:
            d[: nsym * 4].copy_(torch.from_numpy(zipf_tokens(nsym, first_sym).view(np.uint8)))
        else:
            textgen.generate_text(nsym, first=first_sym, kind=cfg["kind"], seed=TEXT_SEED, plant_seed=PLANT_SEED, plant_period=PLANT_PERIOD, dict_flat=flat, dict_offsets=offsets,
                                  device_ptr=d.data_ptr(), stream=stream.cuda_stream)
        return d

    d_text = device_text(first, n)
    # room for every record: sparse configs report ~1 match per 4 KiB; c2 (single-letter words) ~1 per 15 bytes; c5 up to ~1 per token
    cap = {"c2": n // 8, "c5": 2 * n}.get(name, max(1 << 20, n // 512))
    d_matches = torch.empty(cap * 16, dtype=torch.uint8, device="cuda")

    def step_device():
        return m.scan_device(d_text.data_ptr(), n, lead=lead, base=first, d_matches_ptr=d_matches.data_ptr(), capacity=cap, stream=stream.cuda_stream)

    # the path's only collective: one match count per GPU.  It is issued on the scan's stream after every step and left on the
    # device: nobody waits for it inside the timed region, the totals are read once afterwards.
    counts_dev = torch.zeros(1, dtype=torch.int64, device="cuda")
    counts_all = torch.zeros(world, dtype=torch.int64, device="cuda")

    def exchange(local_count):
        if world == 1:
            return
        counts_dev.fill_(local_count)
        dist.all_gather_into_tensor(counts_all, counts_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)  # nvidia-smi needs a moment before its first sample
    t_load = time.time()
    while True:  # warm-up: at least W steps, and long enough for the clock sampler to see the GPU under load
        for _ in range(max(warmup, 1)):
            local_matches = step_device()
            exchange(local_matches)
        more = 0 if (warmup == 0 or time.time() - t_load > 0.25) else 1
        if world > 1:  # every rank must leave the loop after the same number of collectives: rank 0's clock decides
            flag = torch.tensor([more], dtype=torch.int64, device="cuda")
            dist.broadcast(flag, 0)
            more = int(flag.item())
        if not more:
            break
    assert local_matches <= cap, f"record buffer too small: {local_matches} > {cap}"
    st0 = m.stats()
    barrier()
    t_begin = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main_ms, kernel_ms, cands = [], [], 0
    raw = m.stats_into()  # refilled in place below: the scan call synchronises, so every microsecond of host work between two scans is GPU idle time
    ev0.record(stream)
    for _ in range(steps):
        local_matches = step_device()
        exchange(local_matches)
        m.stats_into(raw)
        main_ms.append(raw.main_kernel_ms)
        kernel_ms.append(raw.scan_kernel_ms)
    ev1.record(stream)
    cands = raw.last_nb_candidates
    barrier()
    clocks = sampler.stop(t_begin, time.time())
    st1 = m.stats()
    total_matches = int(counts_all.sum().item()) if world > 1 else local_matches
    ms_local = ev0.elapsed_time(ev1) / max(steps, 1)
    print(f"[{name} rank {rank}] ms/step(local)={ms_local:.3f} kernel_ms={np.mean(kernel_ms):.3f} main_ms={np.mean(main_ms):.3f} fallbacks={st1['fallback_count']} "
          f"stride={st1.get('filter_stride')} matches={local_matches} cands={cands}", file=sys.stderr)
    ms_step = ms_local
    if world > 1:
        t = torch.tensor([ms_local], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item())
    total_bytes = (total_syms if scaling == "strong" else total_syms * world) * width  # owned symbols of all ranks (lead overlaps are overhead)
    value = total_bytes / (ms_step * 1e-3) / 1e9

    # records across ranks: every rank's checksum against the same shard scanned by rank 0 alone (the N-rank run must return what a
    # single GPU returns for the same plan)
    check = None
    if with_check and world > 1:
        mine = torch.tensor([record_hash(torch, d_matches, min(local_matches, cap)) - (1 << 63), local_matches], dtype=torch.int64, device="cuda")
        allh = torch.zeros(world * 2, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(allh, mine)
        if rank == 0:
            allh = allh.cpu().numpy().reshape(world, 2)
            ok, checked = True, 0
            for g, (ga, gb, gl) in enumerate(plan):
                if g == 0:
                    continue
                gn = gb - ga + gl
                if gn * width > (12 << 30):  # keep rank 0's second buffer bounded
                    continue
                d2 = device_text(ga - gl, gn)
                k = m.scan_device(d2.data_ptr(), gn, lead=gl, base=ga - gl, d_matches_ptr=d_matches.data_ptr(), capacity=cap, stream=stream.cuda_stream)
                h = record_hash(torch, d_matches, min(k, cap)) - (1 << 63)
                ok &= (h == int(allh[g, 0])) and (k == int(allh[g, 1]))
                checked += 1
                del d2
            check = {"ranks_checked_against_one_gpu": checked, "identical": bool(ok)}
            assert ok, "records of an N-rank run differ from the single-GPU scan of the same shard"
            local_matches = step_device()  # d_matches holds rank 0's own records again

    # end to end: host (pinned) text in, host (pinned) records out, through the same C-ABI call
    e2e = None
    if with_e2e:
        n_e2e = n if n * width <= (8 << 30) + (1 << 20) else (8 << 30) // width // 4096 * 4096
        h_text = torch.empty(n_e2e * width, dtype=torch.uint8, pin_memory=True)
        h_text.copy_(d_text[: n_e2e * width])
        h_out = torch.empty(cap * 16, dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()
        e2e_steps = max(2, min(steps, 5))
        got = m.scan_host_to_host(h_text.data_ptr(), n_e2e, h_out.data_ptr(), cap, lead=lead, base=first)  # warm-up (allocates the staging buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            got = m.scan_host_to_host(h_text.data_ptr(), n_e2e, h_out.data_ptr(), cap, lead=lead, base=first)
            exchange(got)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        # the same bytes through a plain pinned-host -> device copy, all ranks at once: what the host's PCIe / memory system gives
        d_tmp = torch.empty(n_e2e * width, dtype=torch.uint8, device="cuda")
        d_tmp.copy_(h_text, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            d_tmp.copy_(h_text, non_blocking=True)
        barrier()
        copy_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        del d_tmp
        if world > 1:
            t = torch.tensor([e2e_ms, copy_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms, copy_ms = float(t[0].item()), float(t[1].item())
        assert got <= cap and (n_e2e != n or got == local_matches), (got, local_matches, cap)
        scanned = (n_e2e - lead) * width
        e2e = {"value": scanned * world / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(n_e2e * width), "d2h_bytes_per_step": int(min(got, cap) * 16 + 8),
               "ms_per_step": e2e_ms, "steps": e2e_steps, "bytes_scanned_per_rank": int(scanned), "host_binding": numa,
               "h2d_copy_only_GBps_all_ranks": n_e2e * width * world / (copy_ms * 1e-3) / 1e9}
        del h_text, h_out

    meyer = None
    if cfg.get("meyer_rounds") and rank == 0:
        meyer = meyer_phase(cfg, ac75, torch, m, d_matches, cap, stream, total_syms)

    if rank != 0:
        m.close()
        return None
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
        "metric": "text_GB_per_s_scanned", "value": value, "unit": "GB/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "u8" if width == 1 else "u32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "name": name, "text_bytes_total": text_syms * width, "bytes_per_gpu": owned * width, "engine": st1["engine"], "nb_keywords": st1["nb_keywords"],
                   "nb_states": st1["nb_states"], "l2": "inputs larger than L2 (no flush needed)" if n * width > (256 << 20) else "inputs close to the L2 size: the text streams through L2 between steps",
                   "text_seed": hex(TEXT_SEED), "dict_seed": hex(DICT_SEED), "table_bytes": st1["table_bytes"], "smem_bytes": st1["smem_bytes"],
                   "dictionary_build_s": round(build_s, 2), "finalise_ms": round(st1["finalise_ms"], 1), "filter_stride": st1.get("filter_stride"),
                   "filter_hit_rate": round(st1["filter_fp"], 5), "fallback_count": st1["fallback_count"]},
        "matches_per_s": total_matches / (ms_step * 1e-3), "matches_per_step": total_matches, "candidates_per_step_rank0": cands,
        "kernel_ms": {"main": float(np.mean(main_ms)), "all": float(np.mean(kernel_ms))},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": prof.get(name), "peak_source": peak_src,
                     "kernel": ("filter_scan_s2_kernel" if st1.get("filter_stride") == 2 else "filter_scan_kernel") if st1["engine"] == "filter" else "dfa_scan_kernel (count + event recording) + dfa_emit_events_kernel" if st1.get("dfa_event_scans") else "dfa_scan_kernel (count) + dfa_emit_kernel (walking emit)",
                     "algorithmic_bytes_per_launch": int(algo_bytes)},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(st1["total_kernel_launches"] - st0["total_kernel_launches"]),
    }
    if check is not None:
        line["record_check"] = check
    if meyer is not None:
        line["meyer"] = meyer
    if with_cpu and world == 1:
        oracle, label, kind = reference_oracle(flat, offsets, width, cfg.get("cpu_kind", "reference"))
        sample = min(cfg["cpu_sample"], n * width) // width
        text = host_text(cfg, sample, 0, flat, offsets)
        oracle.count(text[: 1 << 16])  # classic build: lazy construction of the fail links outside the timed region
        oracle.reset_cursor()
        t0 = time.perf_counter()
        cm = oracle.count(text)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": sample * width / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": label,
                                "sample": f"{kind}: acm_match+acm_get_match loop over the first {sample * width >> 20} MiB of the same text, {cm} matches, {dt:.1f} s; host has {host_cores()} cores"}
        oracle.close()
    m.close()
    return line


def meyer_phase(cfg, ac75, torch, m, d_matches, cap, stream, total_syms):
    """BASELINE configs[4]'s incremental part: rounds of +meyer_add keywords, each followed by a scan of the next tenth of the
    stream with the cursor carried across the update (reference aho_corasick_generic_test.c:184-228 interleaves the same way).
    Reports what an update costs (host table update + upload, `finalise_ms` of the scan that finds the machine changed)."""
    rounds, add = cfg["meyer_rounds"], cfg["meyer_add"]
    seg = total_syms // rounds // 4096 * 4096
    text = zipf_tokens(seg * rounds, 0)
    m.reset_cursor()
    update_ms, scan_ms, found, patched = [], [], 0, 0
    for r in range(rounds):
        flat, offsets = ngram_dictionary(add, DICT_SEED + 1 + r)
        t0 = time.perf_counter()
        m.insert_many(flat=flat, offsets=offsets)
        ins_ms = (time.perf_counter() - t0) * 1e3
        before = m.stats()
        t0 = time.perf_counter()
        recs = m.scan(text[r * seg:(r + 1) * seg], base=r * seg, capacity=2 * seg, carry=True)
        wall = (time.perf_counter() - t0) * 1e3
        st = m.stats()
        assert st["finalise_count"] == before["finalise_count"] + 1
        update_ms.append(st["finalise_ms"])
        scan_ms.append(wall - st["finalise_ms"])
        patched += int(st.get("patch_count", 0) - before.get("patch_count", 0))
        found += len(recs)
        del recs
    return {"rounds": rounds, "keywords_added_per_round": add, "tokens_scanned_per_round": seg, "table_update_ms_per_round": [round(x, 2) for x in update_ms],
            "table_update_ms_mean": float(np.mean(update_ms)), "rounds_patched_in_place": patched, "host_insert_ms_last_round": round(ins_ms, 2),
            "scan_wall_ms_mean": float(np.mean(scan_ms)), "matches": found, "nb_keywords_after": m.nb_keywords}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default=None, choices=["strong", "weak"], help="strong: one text of the config's size split over the GPUs; weak: every GPU its own (default: the config's)")
    ap.add_argument("--gib", type=float, default=None, help="GiB of text (per GPU for weak scaling, in total for strong scaling); default: the config's size")
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--option", action="append", default=[], help="key=value passed to acm_b200_set_option")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other configs (extra_configs) and the weak-scaling companion figure")
    ap.add_argument("--with-c4", action="store_true", help="also run configs[3] sharded over the N GPUs (2 GiB per GPU); default at N = 8, its stated shape")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    cfg = dict(CONFIGS[args.config])
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    numa = bind_to_gpu_numa_node(local)  # before torch / CUDA allocate anything pinned
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

    scaling = args.scaling or cfg["scaling"]
    line = measure(args.config, cfg, args, ac75, torch, dist, rank, world, local, scaling, args.steps, args.warmup, not args.no_e2e, not args.no_cpu_baseline, True, numa)
    if not args.no_extra:
        if world > 1 and scaling == "strong":  # the weak-scaling companion: every rank its own text of the config's size
            weak = measure(args.config, cfg, args, ac75, torch, dist, rank, world, local, "weak", max(3, args.steps // 2), max(1, min(args.warmup, 2)), False, False, False, numa)
            if rank == 0:
                line["weak"] = {k: weak[k] for k in ("value", "unit", "ms_per_step", "scaling", "matches_per_step", "kernel_ms")} | {"bytes_per_gpu": weak["config"]["bytes_per_gpu"]}
        if world > 1 and args.config == "c3" and (world == 8 or args.with_c4):
            # configs[3] as stated: 10^6 patterns over 16 GiB sharded across 8 GPUs (2 GiB per GPU; with --with-c4 at any N)
            try:
                c4 = measure("c4", dict(CONFIGS["c4"]), args, ac75, torch, dist, rank, world, local, "weak", max(3, args.steps // 2), max(1, min(args.warmup, 3)), False, False, True, numa)
            except Exception as e:  # must not take the headline line with it (every rank fails or succeeds alike: same code, same sizes)
                c4 = {"config": {"name": "c4"}, "error": f"{type(e).__name__}: {e}"}
            if rank == 0:
                line["extra_configs"] = [c4]
        if world == 1 and args.config == "c3":  # the other configs, driver-run in the same line
            extras = []
            for name in ("c2", "c4s", "c5"):
                try:
                    extras.append(measure(name, dict(CONFIGS[name]), args, ac75, torch, dist, rank, world, local, "weak", max(3, args.steps // 2), max(1, min(args.warmup, 3)),
                                          not args.no_e2e, not args.no_cpu_baseline, False, numa))
                except Exception as e:  # one config failing must not take the headline line with it
                    extras.append({"config": {"name": name}, "error": f"{type(e).__name__}: {e}"})
            line["extra_configs"] = extras
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
