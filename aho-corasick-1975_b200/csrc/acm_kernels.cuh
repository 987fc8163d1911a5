/* acm_kernels.cuh -- the sm_100a kernels of the scan path.  Included by acm_device.cu only.
 *
 * Two engine families, both exact:
 *
 *  DFA engines (byte alphabets).  The text is cut into per-thread chunks; every thread first re-reads the
 *  (max keyword length - 1) symbols before its chunk from state 0 and then walks its chunk through the dense delta table
 *  (one class lookup + one delta lookup per byte).  Pass 1 counts the occurrences of every chunk -- and, in the shared-memory
 *  engine, records where it met an output state -- a device scan turns the counts into offsets, pass 2 expands the recorded
 *  events chunk by chunk (dfa_emit_events_kernel) or, where nothing was recorded, re-walks and writes the records in place
 *  (dfa_emit_kernel) -- so the output is in the reference's emission order without any sort.  delta lives in shared memory
 *  (uint16 entries) or in global memory (uint32 entries, L2 resident).
 *
 *  Filter engine (any alphabet).  A keyword can only END at position p if the q symbols ending at p are the last q symbols
 *  of some keyword (q = min(shortest keyword, 4 bytes / 2 wider symbols)).  Kernel F1 streams the text with coalesced 16-byte
 *  loads, tests every position against a blocked Bloom filter in shared memory (one 32-bit shared load per position), confirms
 *  the survivors in an exact q-gram hash table (L2) and appends the confirmed positions of each warp tile, in position order,
 *  to a candidate list (ballot/popc compaction).  For byte dictionaries whose shortest keyword has at least 4 bytes, F1s does
 *  the same with one shared load per TWO positions (3-byte windows on every second position, a pair table in L2 as second
 *  level) and F1h redoes exactly the few 32 KiB spans whose stages overflowed.  F2 walks the reverse trie leftwards from every
 *  candidate and counts, a device scan gives offsets, F4 walks again and writes the records, longest keyword first.
 */
#pragma once
#include "acm_tables.h"
#include "acm_b200.h"
#include <cuda.h> /* CUtensorMap: the TMA descriptor of the DFA count pass */
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace acm {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xFFFFFFFFu;

/* ------------------------------------------------------------------------------------------------------------------ */
/* Device-wide exclusive scan of uint32 counts into uint64 offsets (three small kernels; the data is a few MB at most). */
/* ------------------------------------------------------------------------------------------------------------------ */
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4; /* per thread */
constexpr int kScanBlock = kScanThreads * kScanItems;

__device__ __forceinline__ uint64_t
block_exclusive_scan (uint64_t v, uint64_t *total, uint64_t *warp_sums /* [32] shared */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint64_t o = __shfl_up_sync (kFull, incl, d);
    if (lane >= d)
      incl += o;
  }
  if (lane == 31)
    warp_sums[warp] = incl;
  __syncthreads ();
  if (warp == 0) {
    uint64_t w = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0, wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint64_t o = __shfl_up_sync (kFull, wi, d);
      if (lane >= d)
        wi += o;
    }
    warp_sums[lane] = wi - w; /* exclusive */
    if (lane == 31)
      *total = wi;
  }
  __syncthreads ();
  uint64_t r = warp_sums[warp] + incl - v;
  __syncthreads ();
  return r;
}

/* Up to kScanSmall counts: the whole scan in one block, 4,096 counts per round; each thread takes 4 consecutive counts, so a warp
 * reads and writes contiguous memory.  A round costs ~3.5 us (measured: 27.7 us for 32 Ki counts, against ~15 us for the three
 * kernels below), so this form is taken for up to two rounds only -- small texts and shards, where three launches cost more than
 * the scan. */
constexpr uint64_t kScanSmall = 8192;
__global__ void __launch_bounds__ (kScanThreads)
scan_small_kernel (const uint32_t *__restrict__ counts, uint64_t n, uint64_t *__restrict__ offsets, uint64_t *__restrict__ grand) {
  __shared__ uint64_t warp_sums[32];
  __shared__ uint64_t total;
  uint64_t carry = 0;
  for (uint64_t start = 0; start < n; start += kScanBlock) {
    const uint64_t base = start + (uint64_t)threadIdx.x * kScanItems;
    uint32_t c[kScanItems];
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
      c[i] = base + i < n ? counts[base + i] : 0;
      v += c[i];
    }
    uint64_t ex = carry + block_exclusive_scan (v, &total, warp_sums);
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
      if (base + i < n)
        offsets[base + i] = ex;
      ex += c[i];
    }
    carry += total; /* the next round writes `total` only after its first barrier */
  }
  if (threadIdx.x == 0)
    *grand = carry;
}

__global__ void __launch_bounds__ (kScanThreads)
scan_block_sums_kernel (const uint32_t *__restrict__ counts, uint64_t n, uint64_t *__restrict__ block_sums) {
  __shared__ uint64_t warp_sums[32];
  __shared__ uint64_t total;
  uint64_t base = (uint64_t)blockIdx.x * kScanBlock + (uint64_t)threadIdx.x * kScanItems, v = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++)
    if (base + i < n)
      v += counts[base + i];
  block_exclusive_scan (v, &total, warp_sums);
  if (threadIdx.x == 0)
    block_sums[blockIdx.x] = total;
}

/* one block: exclusive scan of the block sums in place, grand total to *grand */
__global__ void __launch_bounds__ (kScanThreads)
scan_spine_kernel (uint64_t *__restrict__ block_sums, uint64_t nblocks, uint64_t *__restrict__ grand) {
  __shared__ uint64_t warp_sums[32];
  __shared__ uint64_t total;
  uint64_t carry = 0;
  for (uint64_t start = 0; start < nblocks; start += kScanThreads) {
    uint64_t i = start + threadIdx.x, v = i < nblocks ? block_sums[i] : 0;
    uint64_t ex = block_exclusive_scan (v, &total, warp_sums);
    if (i < nblocks)
      block_sums[i] = carry + ex;
    carry += total;
    __syncthreads ();
  }
  if (threadIdx.x == 0)
    *grand = carry;
}

__global__ void __launch_bounds__ (kScanThreads)
scan_apply_kernel (const uint32_t *__restrict__ counts, uint64_t n, const uint64_t *__restrict__ block_sums, uint64_t *__restrict__ offsets) {
  __shared__ uint64_t warp_sums[32];
  __shared__ uint64_t total;
  uint64_t base = (uint64_t)blockIdx.x * kScanBlock + (uint64_t)threadIdx.x * kScanItems, v = 0;
  uint32_t c[kScanItems];
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    c[i] = base + i < n ? counts[base + i] : 0;
    v += c[i];
  }
  uint64_t ex = block_exclusive_scan (v, &total, warp_sums) + block_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    if (base + i < n)
      offsets[base + i] = ex;
    ex += c[i];
  }
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* DFA engines                                                                                                         */
/* ------------------------------------------------------------------------------------------------------------------ */
struct DfaParams {
  const uint8_t *text;
  uint64_t n;     /* symbols */
  uint64_t lead;  /* occurrences ending before `lead` are not reported */
  uint64_t base;  /* added to reported end positions */
  uint64_t chunk; /* symbols per chunk, multiple of 16 */
  uint64_t nchunks;
  uint32_t warm;       /* symbols re-read before a chunk = max keyword length - 1 */
  uint32_t init_state; /* dfa state chunk 0 starts from (carried cursor) */
  const void *delta;   /* [nb_states][K] */
  uint32_t K, nb_states, out_threshold;
  const uint32_t *out_offsets;
  const acm_output *out_entries;
  uint32_t nb_out_states;        /* states >= out_threshold */
  uint32_t counts_in_smem;       /* pass 1 keeps a uint16 records-per-state table in shared memory (small dictionaries) */
  uint64_t first_chunk;          /* the walking kernels start at this chunk (the chunks before it were done by the TMA-staged pass) */
  uint64_t tma_chunks;           /* chunks [0, tma_chunks) lie wholly inside the text: rows of the TMA tensor */
  uint32_t *chunk_counts;        /* pass 1 out */
  const uint64_t *chunk_offsets; /* pass 2 in */
  ACMB200Match *matches;
  uint64_t capacity;
  /* pass 1 can record WHERE it met an output state (position inside the chunk | output state << 16), so that pass 2 expands the
   * recorded events instead of walking the text a second time (shared-memory engine, chunks of at most 65,536 symbols) */
  uint32_t *events;         /* [nchunks][events_per_chunk] */
  uint32_t *chunk_events;   /* [nchunks] events met (may exceed events_per_chunk: then *events_overflow is set and pass 2 walks) */
  uint32_t events_per_chunk;
  uint32_t *events_overflow;
  uint8_t class_of_byte[256];
};

template <typename Entry, bool kShared, bool kEmit, bool kEvents = false>
__global__ void __launch_bounds__ (kShared ? 1024 : 256, kShared ? 1 : 2)
dfa_scan_kernel (const __grid_constant__ DfaParams p) {
  extern __shared__ __align__ (16) unsigned char smem[];
  uint8_t *s_class = smem;
  Entry *s_delta = reinterpret_cast<Entry *> (smem + 256);
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    s_class[i] = p.class_of_byte[i];
  if (kShared) {
    /* table bytes rounded up to 16 by the host */
    const uint4 *src = reinterpret_cast<const uint4 *> (p.delta);
    uint4 *dst = reinterpret_cast<uint4 *> (s_delta);
    const uint32_t vecs = (uint32_t)(((uint64_t)p.nb_states * p.K * sizeof (Entry) + 15) / 16);
    for (uint32_t i = threadIdx.x; i < vecs; i += blockDim.x)
      dst[i] = src[i];
  }
  /* records per output state, next to the table, when the host found room for it (pass 1 only needs the count) */
  uint16_t *s_counts = reinterpret_cast<uint16_t *> (smem + 256 + (kShared ? ((size_t)p.nb_states * p.K * sizeof (Entry) + 15) / 16 * 16 : 0));
  if (!kEmit && p.counts_in_smem)
    for (uint32_t i = threadIdx.x; i < p.nb_out_states; i += blockDim.x)
      s_counts[i] = (uint16_t)(p.out_offsets[i + 1] - p.out_offsets[i]);
  __syncthreads ();
  const Entry *__restrict__ delta = kShared ? s_delta : reinterpret_cast<const Entry *> (p.delta);
  const uint32_t K = p.K, thr = p.out_threshold;
  const bool smem_counts = !kEmit && p.counts_in_smem;

  for (uint64_t c = p.first_chunk + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < p.nchunks; c += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t start = c * p.chunk;
    const uint64_t end = min (p.n, start + p.chunk);
    const uint64_t report_from = max (start, p.lead);
    /* chunks that start within `warm` symbols of the text start re-read from symbol 0, where the true state is the carried one */
    const bool from_origin = start <= p.warm;
    uint32_t state = from_origin ? p.init_state : 0;
    uint64_t pos = from_origin ? 0 : (start - p.warm) & ~(uint64_t)15;
    uint32_t count = 0, nev = 0;
    uint64_t out = kEmit ? p.chunk_offsets[c] : 0;
    uint32_t *my_events = kEvents ? p.events + c * p.events_per_chunk : nullptr;

    /* positions are walked relative to the chunk start, in 32 bits (negative during the warm-up): the per-byte test is one
     * compare, and the recorded event needs no 64-bit arithmetic */
    const int32_t report_rel = (int32_t)min (report_from - start, p.chunk); /* a chunk wholly before `lead` never reports */
    auto step = [&] (uint32_t byte, int32_t rel) {
      state = delta[state * K + s_class[byte]];
      if (state >= thr && rel >= report_rel) {
        const uint32_t o = state - thr;
        const uint64_t at = start + (int64_t)rel; /* only the emitting variant uses it */
        if (kEvents) {
          if (nev < p.events_per_chunk)
            *my_events = (uint32_t)rel | (o << 16);
          my_events++; /* a running pointer: no index arithmetic per event */
          nev++;
        }
        if (smem_counts) {
          count += s_counts[o];
          return;
        }
        const uint32_t lo = p.out_offsets[o], hi = p.out_offsets[o + 1];
        if (kEmit) {
          for (uint32_t j = lo; j < hi; j++, out++)
            if (out < p.capacity) {
              const acm_output e = p.out_entries[j];
              p.matches[out] = ACMB200Match{ p.base + at, e.keyword, e.length };
            }
        } else
          count += hi - lo;
      }
    };
    /* warm-up + chunk, 16 bytes at a time while a whole vector is inside the text.  A thread's loads are a dependent chain of
     * DRAM round trips (nothing else of its chunk is in flight), so the next vector is requested before the current one is walked
     * and the line after that is pulled into L2. */
    /* 32 bytes (one sector) per round while two whole vectors are inside the text, the next round's pair requested before this
     * round is walked; then at most one single vector */
    auto round16 = [&] () {
      const uint4 v = *reinterpret_cast<const uint4 *> (p.text + pos);
      const uint32_t w[4] = { v.x, v.y, v.z, v.w };
      const int32_t rel0 = (int32_t)((int64_t)pos - (int64_t)start);
#pragma unroll
      for (int i = 0; i < 16; i++)
        step ((w[i >> 2] >> (8 * (i & 3))) & 0xFFu, rel0 + i);
      pos += 16;
    };
    if ((pos & 31) && pos + 16 <= end)
      round16 (); /* chunks start on 16-byte boundaries: get onto a sector boundary */
    bool have = pos + 32 <= end;
    uint4 na = make_uint4 (0, 0, 0, 0), nb = na;
    if (have) {
      na = *reinterpret_cast<const uint4 *> (p.text + pos);
      nb = *reinterpret_cast<const uint4 *> (p.text + pos + 16);
    }
    while (have) {
      const uint4 va = na, vb = nb;
      const int32_t rel0 = (int32_t)((int64_t)pos - (int64_t)start);
      pos += 32;
      have = pos + 32 <= end;
      if (have) {
        na = *reinterpret_cast<const uint4 *> (p.text + pos);
        nb = *reinterpret_cast<const uint4 *> (p.text + pos + 16);
        if ((pos & 127) == 0 && pos + 256 <= end)
          asm volatile ("prefetch.global.L2 [%0];" ::"l"(p.text + pos + 128));
      }
      const uint32_t w[8] = { va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w };
#pragma unroll
      for (int i = 0; i < 32; i++)
        step ((w[i >> 2] >> (8 * (i & 3))) & 0xFFu, rel0 + i);
    }
    while (pos + 16 <= end)
      round16 ();
    for (; pos < end; pos++)
      step (p.text[pos], (int32_t)((int64_t)pos - (int64_t)start));
    if (!kEmit)
      p.chunk_counts[c] = count;
    if (kEvents) {
      p.chunk_events[c] = nev;
      if (nev > p.events_per_chunk)
        atomicExch (p.events_overflow, 1u);
    }
  }
}

/* Pass 1 of the shared-memory DFA engine with the text staged through shared memory by the TMA unit.
 *
 * The text is a 2-D tensor of bytes: row = chunk, `chunk` bytes per row (acm_device.cu encodes the tensor map per scan).  A warp owns
 * 32 neighbouring chunks; one cp.async.bulk.tensor.2d per stage brings the next 32 bytes of ALL 32 chunks (a 32 x 32-byte box: one
 * sector per row, 1 KB) into the warp's stage buffer and completes on the warp's mbarrier, two stages in flight.  Every lane then
 * takes its row with two 16-byte shared loads.  What the per-thread loads of dfa_scan_kernel cost -- 32 lines touched per warp-wide
 * instruction, address arithmetic, a dependent chain of global round trips per lane -- is one instruction of one lane here.
 * The warm-up (max keyword length - 1 bytes before the chunk) is the tail of the row above: the same box one row up.
 * The walk itself records events WITHOUT branching (an output state is met every ~15 bytes of config 2's text, so a branch on it
 * splits the warp at almost every step): predicated store through a clamped running pointer, predicated count. */
constexpr uint32_t kTmaStageBytes = 32;                 /* per chunk and stage */
constexpr uint32_t kTmaWarpBytes = 2 * 32 * kTmaStageBytes; /* two stages of 32 rows */
constexpr uint32_t kTmaCtaBytes = 32 * kTmaWarpBytes + 32 * 2 * 8; /* stage buffers of 32 warps + their mbarriers */

__device__ __forceinline__ uint32_t
smem_u32 (const void *p) {
  return (uint32_t)__cvta_generic_to_shared (p);
}

template <bool kEvents>
__global__ void __launch_bounds__ (1024, 1)
dfa_scan_tma_kernel (const __grid_constant__ DfaParams p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__ (128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *stage_buf = smem + (size_t)warp * kTmaWarpBytes;
  uint64_t *mbar = reinterpret_cast<uint64_t *> (smem + 32 * kTmaWarpBytes) + 2 * warp;
  uint8_t *s_class = smem + kTmaCtaBytes;
  uint16_t *s_delta = reinterpret_cast<uint16_t *> (smem + kTmaCtaBytes + 256);
  const size_t delta_bytes = ((size_t)p.nb_states * p.K * 2 + 15) / 16 * 16;
  uint16_t *s_counts = reinterpret_cast<uint16_t *> (smem + kTmaCtaBytes + 256 + delta_bytes);
  if (lane == 0) {
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32 (&mbar[0])));
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32 (&mbar[1])));
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    s_class[i] = p.class_of_byte[i];
  {
    const uint4 *src = reinterpret_cast<const uint4 *> (p.delta);
    uint4 *dst = reinterpret_cast<uint4 *> (s_delta);
    for (uint32_t i = threadIdx.x; i < delta_bytes / 16; i += blockDim.x)
      dst[i] = src[i];
  }
  for (uint32_t i = threadIdx.x; i <= p.nb_out_states; i += blockDim.x)
    s_counts[i] = i < p.nb_out_states ? (uint16_t)(p.out_offsets[i + 1] - p.out_offsets[i]) : (uint16_t)0;
  __syncthreads ();

  const uint32_t K = p.K, thr = p.out_threshold;
  const uint32_t chunk = (uint32_t)p.chunk, warm_pad = (p.warm + kTmaStageBytes - 1) / kTmaStageBytes * kTmaStageBytes;
  const uint32_t warm_stages = warm_pad / kTmaStageBytes, stages = warm_stages + chunk / kTmaStageBytes;
  uint32_t g = 0; /* stages issued so far by this warp: buffer g & 1, mbarrier phase (g >> 1) & 1 */
  /* lane 0 asks the TMA unit for stage k of the warp's rows [row0, row0 + 32): the warm-up stages read the row above */
  auto issue = [&] (uint32_t k, uint64_t row0, uint32_t slot) {
    if (lane == 0) {
      const int32_t x = k < warm_stages ? (int32_t)(chunk - warm_pad + k * kTmaStageBytes) : (int32_t)((k - warm_stages) * kTmaStageBytes);
      const int32_t y = (int32_t)row0 - (k < warm_stages ? 1 : 0);
      const uint32_t bar = smem_u32 (&mbar[slot]), dst = smem_u32 (stage_buf + slot * 32 * kTmaStageBytes);
      asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(32u * kTmaStageBytes) : "memory");
      asm volatile ("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(&tmap), "r"(bar), "r"(x), "r"(y)
                    : "memory");
    }
  };

  for (uint64_t row0 = ((uint64_t)blockIdx.x * 32 + warp) * 32; row0 < p.tma_chunks; row0 += (uint64_t)gridDim.x * 32 * 32) {
    const uint64_t c = row0 + lane;
    const bool have = c < p.tma_chunks; /* rows beyond the tensor read as zeros and are not reported */
    const uint64_t start = c * p.chunk;
    /* chunk 0 starts from the carried cursor's state and has no warm-up; the others re-read the tail of the row above from state 0 */
    uint32_t state = c == 0 ? p.init_state : 0;
    const int32_t report_rel = have ? (int32_t)min (p.lead > start ? p.lead - start : (uint64_t)0, p.chunk) : (int32_t)chunk;
    uint32_t count = 0;
    uint32_t *const ev_first = kEvents ? p.events + c * p.events_per_chunk : nullptr;
    const uint32_t ev_cap_m1 = kEvents ? p.events_per_chunk - 1 : 0; /* a chunk with more events keeps overwriting its last slot */
    const uint32_t nb_out = p.nb_out_states; /* s_counts[nb_out] == 0 */
    const uint32_t ev_slot0 = (uint32_t)(c * p.events_per_chunk); /* (the host uses this kernel only while the event buffer has fewer than 2^32 slots) */
    uint32_t nev = 0, ev_slot = ev_slot0;

    issue (0, row0, g & 1);
    if (stages > 1)
      issue (1, row0, (g + 1) & 1);
    for (uint32_t k = 0; k < stages; k++, g++) {
      const uint32_t slot = g & 1, phase = (g >> 1) & 1;
      { /* wait for the stage */
        const uint32_t bar = smem_u32 (&mbar[slot]);
        uint32_t done = 0;
        while (!done)
          asm volatile ("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
      }
      const uint4 *row = reinterpret_cast<const uint4 *> (stage_buf + slot * 32 * kTmaStageBytes + lane * kTmaStageBytes);
      const uint4 va = row[0], vb = row[1];
      /* the shared-memory pipe is this kernel's bottleneck, so the two loads above can sit in its queue for a long time; the TMA unit
       * writes through the async proxy and does not know about them: a proxy fence orders this lane's reads before the refill that
       * lane 0 asks for after the warp barrier (without it about one stage in 10,000 was overwritten before it had been read) */
      asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp (); /* every lane has its 32 bytes: the buffer may be refilled */
      if (k + 2 < stages)
        issue (k + 2, row0, slot);
      const uint32_t w[8] = { va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w };
      if (k < warm_stages) { /* warm-up: walked, never reported */
        if (c != 0) {
#pragma unroll
          for (int i = 0; i < 32; i++)
            state = s_delta[state * K + s_class[(w[i >> 2] >> (8 * (i & 3))) & 0xFFu]];
        }
      } else {
        const int32_t rel0 = (int32_t)((k - warm_stages) * kTmaStageBytes);
        if (rel0 >= report_rel && (!kEvents || nev + 32 <= p.events_per_chunk)) {
          /* usual stage: every byte reportable, room for 32 more events.  Straight-line: the records of the state come from a
           * table that has a zero entry for "no output" (index clamped, no predicate), the event goes out through ONE predicated
           * store at a running 32-bit slot index */
#pragma unroll
          for (int i = 0; i < 32; i++) {
            state = s_delta[state * K + s_class[(w[i >> 2] >> (8 * (i & 3))) & 0xFFu]];
            const uint32_t o = state - thr; /* wraps for states without outputs */
            count += s_counts[min (o, nb_out)];
            if (kEvents) {
              if (state >= thr)
                p.events[ev_slot] = (o << 16) + (uint32_t)rel0 + (uint32_t)i;
              ev_slot += state >= thr;
            }
          }
          if (kEvents)
            nev = ev_slot - ev_slot0;
        } else { /* the stage that holds the caller's lead: per-byte test */
#pragma unroll
          for (int i = 0; i < 32; i++) {
            state = s_delta[state * K + s_class[(w[i >> 2] >> (8 * (i & 3))) & 0xFFu]];
            if (state >= thr && rel0 + i >= report_rel) {
              const uint32_t o = state - thr;
              count += s_counts[o];
              if (kEvents) {
                ev_first[min (nev, ev_cap_m1)] = (uint32_t)(rel0 + i) | (o << 16);
                nev++;
                ev_slot = ev_slot0 + min (nev, ev_cap_m1);
              }
            }
          }
        }
      }
    }
    if (have) {
      p.chunk_counts[c] = count;
      if (kEvents) {
        p.chunk_events[c] = nev;
        if (nev > p.events_per_chunk)
          atomicExch (p.events_overflow, 1u);
      }
    }
    __syncwarp ();
  }
}

/* Lean form of the TMA-staged pass 1 (the usual pass 1 of `dfa_smem` when events are recorded): the walk records its events and
 * nothing else.  The records of a chunk are counted afterwards from its event list (dfa_count_events_kernel), so the per-byte
 * work loses the third shared-memory lookup (records per state), the index clamp and the add: 9 instructions per byte instead
 * of 22 --
 *   PRMT (byte), LDS.U8 (2 x class), IMAD (row address), LDS.U16 (next state), IMAD (event word), IADD (position), ISETP,
 *   and under the predicate one STG and one 32-bit pointer increment.
 * The tricks: the class table holds 2 x class and the row stride is 2 K bytes, so one IMAD yields the byte address of the next
 * delta entry; the tables are addressed as [register + uniform base + immediate]; the event pointer is a 32-bit low word next to
 * a constant high word (the host guarantees that a chunk's event list does not cross a 4 GiB boundary), so a recorded event
 * advances it with ONE predicated add; the event word (output state index << 16 | position) is state * 65536 + a running
 * position that already carries -threshold << 16. */
__global__ void __launch_bounds__ (1024, 1)
dfa_scan_tma_lean_kernel (const __grid_constant__ DfaParams p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__ (128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *stage_buf = smem + (size_t)warp * kTmaWarpBytes;
  uint64_t *mbar = reinterpret_cast<uint64_t *> (smem + 32 * kTmaWarpBytes) + 2 * warp;
  uint8_t *s_class2 = smem + kTmaCtaBytes;
  const unsigned char *s_delta8 = smem + kTmaCtaBytes + 256;
  const size_t delta_bytes = ((size_t)p.nb_states * p.K * 2 + 15) / 16 * 16;
  if (lane == 0) {
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32 (&mbar[0])));
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32 (&mbar[1])));
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    s_class2[i] = (uint8_t)(2u * p.class_of_byte[i]); /* the host takes this kernel only while 2 (K - 1) fits a byte */
  {
    const uint4 *src = reinterpret_cast<const uint4 *> (p.delta);
    uint4 *dst = reinterpret_cast<uint4 *> (smem + kTmaCtaBytes + 256);
    for (uint32_t i = threadIdx.x; i < delta_bytes / 16; i += blockDim.x)
      dst[i] = src[i];
  }
  __syncthreads ();

  const uint32_t K2 = 2u * p.K, thr = p.out_threshold;
  const uint32_t chunk = (uint32_t)p.chunk, warm_pad = (p.warm + kTmaStageBytes - 1) / kTmaStageBytes * kTmaStageBytes;
  const uint32_t warm_stages = warm_pad / kTmaStageBytes, stages = warm_stages + chunk / kTmaStageBytes;
  const uint32_t ev_cap4 = p.events_per_chunk * 4u;
  uint32_t g = 0;
  auto issue = [&] (uint32_t k, uint64_t row0, uint32_t slot) {
    if (lane == 0) {
      const int32_t x = k < warm_stages ? (int32_t)(chunk - warm_pad + k * kTmaStageBytes) : (int32_t)((k - warm_stages) * kTmaStageBytes);
      const int32_t y = (int32_t)row0 - (k < warm_stages ? 1 : 0);
      const uint32_t bar = smem_u32 (&mbar[slot]), dst = smem_u32 (stage_buf + slot * 32 * kTmaStageBytes);
      asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(32u * kTmaStageBytes) : "memory");
      asm volatile ("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(&tmap), "r"(bar), "r"(x), "r"(y)
                    : "memory");
    }
  };
  /* the two tables by their shared-space addresses, held in uniform registers (a warp reduction is how a value gets there):
   * the lookups are LDS [register + uniform register], no address add per byte */
  const uint32_t class_base = __reduce_max_sync (kFull, smem_u32 (s_class2)), delta_base = __reduce_max_sync (kFull, smem_u32 (s_delta8));
  auto next_state = [&] (uint32_t state, uint32_t word, int b) -> uint32_t {
    uint32_t c2, next, at;
    asm ("ld.shared.u8 %0, [%1];" : "=r"(c2) : "r"(__byte_perm (word, 0u, 0x4440 + b) + class_base));
    asm ("mad.lo.u32 %0, %1, %2, %3;" : "=r"(at) : "r"(state), "r"(K2), "r"(c2)); /* opaque: ONE multiply-add on the state -> state chain */
    asm ("ld.shared.u16 %0, [%1];" : "=r"(next) : "r"(at + delta_base));
    return next;
  };

  for (uint64_t row0 = ((uint64_t)blockIdx.x * 32 + warp) * 32; row0 < p.tma_chunks; row0 += (uint64_t)gridDim.x * 32 * 32) {
    const uint64_t c = row0 + lane;
    const bool have = c < p.tma_chunks; /* rows beyond the tensor read as zeros and are not reported */
    const uint64_t start = c * p.chunk;
    uint32_t state = c == 0 ? p.init_state : 0;
    const int32_t report_rel = have ? (int32_t)min (p.lead > start ? p.lead - start : (uint64_t)0, p.chunk) : (int32_t)chunk;
    uint64_t ev_ptr = reinterpret_cast<uint64_t> (p.events + c * p.events_per_chunk); /* only its low word ever changes */
    const uint32_t ev_lo0 = (uint32_t)ev_ptr;
    uint32_t dropped = 0; /* events met while the list was full (then pass 2 walks the text instead) */
#define ev_lo ((uint32_t)ev_ptr)

    issue (0, row0, g & 1);
    if (stages > 1)
      issue (1, row0, (g + 1) & 1);
    for (uint32_t k = 0; k < stages; k++, g++) {
      const uint32_t slot = g & 1, phase = (g >> 1) & 1;
      {
        const uint32_t bar = smem_u32 (&mbar[slot]);
        uint32_t done = 0;
        while (!done)
          asm volatile ("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
      }
      const uint4 *row = reinterpret_cast<const uint4 *> (stage_buf + slot * 32 * kTmaStageBytes + lane * kTmaStageBytes);
      const uint4 va = row[0], vb = row[1];
      asm volatile ("fence.proxy.async.shared::cta;" ::: "memory"); /* this lane's reads before the refill (see dfa_scan_tma_kernel) */
      __syncwarp ();
      if (k + 2 < stages)
        issue (k + 2, row0, slot);
      const uint32_t w[8] = { va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w };
      if (k < warm_stages) { /* warm-up: walked, never reported */
        if (c != 0) {
#pragma unroll
          for (int i = 0; i < 32; i++)
            state = next_state (state, w[i >> 2], i & 3);
        }
      } else {
        const int32_t rel0 = (int32_t)((k - warm_stages) * kTmaStageBytes);
        if (rel0 >= report_rel && (ev_lo - ev_lo0) + 32u * 4u <= ev_cap4) {
          /* usual stage: every byte reportable, room for 32 more events: straight-line, one predicated store per byte */
          uint32_t relc = (uint32_t)rel0 - (thr << 16);
#pragma unroll
          for (int i = 0; i < 32; i++) {
            state = next_state (state, w[i >> 2], i & 3);
            const uint32_t val = state * 65536u + relc; /* (state - thr) << 16 | position */
            relc++;
            asm volatile ("{\n\t.reg .pred p;\n\t.reg .b32 lo, hi;\n\tsetp.ge.u32 p, %1, %2;\n\t@p st.global.u32 [%0], %3;\n\tmov.b64 {lo, hi}, %0;\n\t@p add.u32 lo, lo, 4;\n\tmov.b64 %0, {lo, hi};\n\t}"
                          : "+l"(ev_ptr)
                          : "r"(state), "r"(thr), "r"(val)
                          : "memory");
          }
        } else { /* the stage that holds the caller's lead, or a nearly full event list: per-byte tests */
#pragma unroll
          for (int i = 0; i < 32; i++) {
            state = next_state (state, w[i >> 2], i & 3);
            if (state >= thr && rel0 + i >= report_rel) {
              if (ev_lo - ev_lo0 < ev_cap4) {
                *reinterpret_cast<uint32_t *> (ev_ptr) = (uint32_t)(rel0 + i) | ((state - thr) << 16);
                ev_ptr += 4; /* no carry: the list does not cross a 4 GiB boundary */
              } else
                dropped++;
            }
          }
        }
      }
    }
    if (have) {
      p.chunk_events[c] = (ev_lo - ev_lo0) / 4u + dropped;
      if (dropped)
        atomicExch (p.events_overflow, 1u);
    }
#undef ev_lo
    __syncwarp ();
  }
}

/* Records per chunk from the recorded events (after the lean pass 1): a warp per chunk sums the sizes of the output sets of the
 * chunk's events.  A chunk whose list overflowed has set *events_overflow: the host then counts by walking. */
__global__ void __launch_bounds__ (256)
dfa_count_events_kernel (const __grid_constant__ DfaParams p) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp_global = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t c = warp_global; c < p.nchunks; c += nwarps) {
    const uint32_t nev = min (p.chunk_events[c], p.events_per_chunk);
    const uint32_t *ev = p.events + c * p.events_per_chunk;
    uint32_t sum = 0;
    for (uint32_t i = lane; i < nev; i += 32) {
      const uint32_t o = ev[i] >> 16;
      sum += p.out_offsets[o + 1] - p.out_offsets[o];
    }
    sum = __reduce_add_sync (kFull, sum);
    if (lane == 0)
      p.chunk_counts[c] = sum;
  }
}

/* Pass 2 of the shared-memory DFA engine when pass 1 recorded its events: no walk, no table.  A warp takes one chunk at a time and
 * expands 32 of its events per step: records per event from the CSR offsets, a warp scan for the positions, and stores that are
 * contiguous across the warp (the records of a chunk are contiguous in the output, in position order, longest keyword first). */
__global__ void __launch_bounds__ (256)
dfa_emit_events_kernel (const __grid_constant__ DfaParams p) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp_global = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t c = warp_global; c < p.nchunks; c += nwarps) {
    const uint32_t nev = p.chunk_events[c];
    if (!nev)
      continue;
    const uint32_t *ev = p.events + c * p.events_per_chunk;
    const uint64_t chunk_start = c * p.chunk;
    uint64_t out = p.chunk_offsets[c];
    for (uint32_t i0 = 0; i0 < nev; i0 += 32) {
      uint32_t lo = 0, n = 0, rel = 0;
      if (i0 + lane < nev) {
        const uint32_t e = ev[i0 + lane], o = e >> 16;
        rel = e & 0xFFFFu;
        lo = p.out_offsets[o];
        n = p.out_offsets[o + 1] - lo;
      }
      uint32_t incl = n;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync (kFull, incl, d);
        if (lane >= d)
          incl += v;
      }
      const uint32_t total = __shfl_sync (kFull, incl, 31);
      uint64_t at = out + (incl - n);
      for (uint32_t j = 0; j < n; j++, at++)
        if (at < p.capacity) {
          const acm_output e = p.out_entries[lo + j];
          p.matches[at] = ACMB200Match{ p.base + chunk_start + rel, e.keyword, e.length };
        }
      out += total;
    }
  }
}

/* Pass 2 of the DFA engines, warp-cooperative.  The 32 lanes of a warp walk 32 neighbouring chunks in lockstep.  Whenever a lane
 * reaches a state with outputs it appends an event (lane, position, state) to the warp's queue in shared memory (ballot/popc
 * compaction); as soon as 32 events are queued, ALL lanes expand one event each into its records.  The output cursor of every
 * chunk lives in shared memory; events of the same chunk inside one batch are ordered with match_any, so each chunk's records
 * stay in position order and the whole output stays in the reference's emission order. */
constexpr int kEmitQueue = 64; /* events per warp */
struct EmitWarpState {
  unsigned long long cursor[32]; /* next record index of each lane's chunk */
  uint2 queue[kEmitQueue];       /* x = position relative to the warp's first chunk, y = lane << 27 | state */
  uint32_t batch_records[32];    /* records of each event of the batch being expanded */
};

/* Expands the first `count` (<= 32) queued events of a warp, one per lane, into records; returns the new queue length. */
__device__ __noinline__ uint32_t
emit_drain (EmitWarpState &ws, const DfaParams &p, uint32_t thr, uint64_t warp_origin, uint32_t queued, uint32_t count) {
  const int lane = threadIdx.x & 31;
  const uint32_t lanes_below = (1u << lane) - 1u;

      const bool mine = (uint32_t)lane < count;
      uint32_t L = 32u + lane, lo = 0, n = 0;
      uint64_t at_pos = 0;
      if (mine) {
        const uint2 ev = ws.queue[lane];
        L = ev.y >> 27;
        const uint32_t o = (ev.y & 0x07FFFFFFu) - thr;
        lo = p.out_offsets[o];
        n = p.out_offsets[o + 1] - lo;
        at_pos = warp_origin + ev.x;
      }
      /* events of the same chunk in this batch: earlier ones (lower queue index) come first */
      ws.batch_records[lane] = n;
      const uint32_t same = __match_any_sync (kFull, L);
      const bool last_of_chunk = mine && (same >> lane) == 1u;
      __syncwarp ();
      uint32_t before = 0;
      for (uint32_t earlier = same & lanes_below; earlier; earlier &= earlier - 1)
        before += ws.batch_records[__ffs (earlier) - 1];
      unsigned long long out = mine ? ws.cursor[L] + before : 0;
      __syncwarp ();
      if (last_of_chunk)
        ws.cursor[L] = out + n;
      for (uint32_t j = 0; j < n; j++, out++)
        if (out < p.capacity) {
          const acm_output e = p.out_entries[lo + j];
          p.matches[out] = ACMB200Match{ p.base + at_pos, e.keyword, e.length };
        }
      __syncwarp ();
      /* shift the rest of the queue to the front */
      const uint32_t rest = queued - count;
      uint2 moved = make_uint2 (0, 0);
      if ((uint32_t)lane < rest)
        moved = ws.queue[count + lane];
      __syncwarp ();
      if ((uint32_t)lane < rest)
        ws.queue[lane] = moved;
      __syncwarp ();
      return rest;
    }

template <typename Entry, bool kShared>
__global__ void __launch_bounds__ (kShared ? 1024 : 256, kShared ? 1 : 2)
dfa_emit_kernel (const __grid_constant__ DfaParams p) {
  extern __shared__ __align__ (16) unsigned char smem[];
  uint8_t *s_class = smem;
  EmitWarpState *s_emit = reinterpret_cast<EmitWarpState *> (smem + 256);
  Entry *s_delta = reinterpret_cast<Entry *> (smem + 256 + sizeof (EmitWarpState) * (blockDim.x >> 5));
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    s_class[i] = p.class_of_byte[i];
  if (kShared) {
    const uint4 *src = reinterpret_cast<const uint4 *> (p.delta);
    uint4 *dst = reinterpret_cast<uint4 *> (s_delta);
    const uint32_t vecs = (uint32_t)(((uint64_t)p.nb_states * p.K * sizeof (Entry) + 15) / 16);
    for (uint32_t i = threadIdx.x; i < vecs; i += blockDim.x)
      dst[i] = src[i];
  }
  __syncthreads ();
  const Entry *__restrict__ delta = kShared ? s_delta : reinterpret_cast<const Entry *> (p.delta);
  const uint32_t K = p.K, thr = p.out_threshold;
  const int lane = threadIdx.x & 31;
  const uint32_t lanes_below = (1u << lane) - 1u;
  EmitWarpState &ws = s_emit[threadIdx.x >> 5];
  const uint64_t warp_global = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;

  for (uint64_t c0 = warp_global * 32; c0 < p.nchunks; c0 += nwarps * 32) {
    const uint64_t c = c0 + lane;
    const bool have = c < p.nchunks;
    const uint64_t warp_origin = c0 * p.chunk;
    const uint64_t start = have ? c * p.chunk : p.n;
    const uint64_t end = have ? min (p.n, start + p.chunk) : p.n;
    const uint64_t report_from = max (start, p.lead);
    const bool from_origin = start <= p.warm;
    uint32_t state = from_origin ? p.init_state : 0;
    uint64_t pos = !have ? p.n : (from_origin ? 0 : (start - p.warm) & ~(uint64_t)15);
    ws.cursor[lane] = have ? p.chunk_offsets[c] : 0;
    uint32_t queued = 0; /* warp-uniform */
    __syncwarp ();

    auto drain = [&] (uint32_t count) { queued = emit_drain (ws, p, thr, warp_origin, queued, count); };

    /* the next vector of the lane's chunk is requested one round ahead (see dfa_scan_kernel) */
    uint4 nxt = make_uint4 (0, 0, 0, 0);
    if (pos + 16 <= end)
      nxt = *reinterpret_cast<const uint4 *> (p.text + pos);
    while (__any_sync (kFull, pos < end)) {
      /* next 16 bytes of this lane's walk (fewer at the end of its chunk) */
      uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0, nvalid = 0;
      if (pos + 16 <= end) {
        w0 = nxt.x, w1 = nxt.y, w2 = nxt.z, w3 = nxt.w;
        nvalid = 16;
        if (pos + 32 <= end) {
          nxt = *reinterpret_cast<const uint4 *> (p.text + pos + 16);
          if (((pos + 16) & 127) == 0 && pos + 272 <= end)
            asm volatile ("prefetch.global.L2 [%0];" ::"l"(p.text + pos + 144));
        }
      } else if (pos < end) {
        nvalid = (uint32_t)(end - pos);
#pragma unroll
        for (int i = 0; i < 16; i++) { /* compile-time i: the words stay in registers */
          const uint32_t b = (uint32_t)i < nvalid ? (uint32_t)p.text[pos + i] << (8 * (i & 3)) : 0u;
          if ((i >> 2) == 0)
            w0 |= b;
          else if ((i >> 2) == 1)
            w1 |= b;
          else if ((i >> 2) == 2)
            w2 |= b;
          else
            w3 |= b;
        }
      }
      /* bytes before `first_reported` belong to the warm-up (or the caller's lead): walked, never reported */
      const uint32_t first_reported = report_from > pos ? (uint32_t)min ((uint64_t)16, report_from - pos) : 0u;
      const uint32_t rel0 = (uint32_t)(pos - warp_origin), tag = (uint32_t)lane << 27;
      auto enqueue = [&] (uint32_t evmask, bool event, int i) {
        if (event)
          ws.queue[queued + __popc (evmask & lanes_below)] = make_uint2 (rel0 + i, tag | state);
        queued += __popc (evmask);
        __syncwarp ();
        if (queued >= 32)
          drain (32);
      };
      if (__all_sync (kFull, nvalid == 16 && first_reported == 0)) {
        /* usual case, warp-uniform: every lane has 16 reportable bytes -- no per-byte validity logic */
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const uint32_t word = (i >> 2) == 0 ? w0 : ((i >> 2) == 1 ? w1 : ((i >> 2) == 2 ? w2 : w3));
          state = delta[state * K + s_class[(word >> (8 * (i & 3))) & 0xFFu]];
          const bool event = state >= thr;
          const uint32_t evmask = __ballot_sync (kFull, event);
          if (evmask)
            enqueue (evmask, event, i);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const uint32_t word = (i >> 2) == 0 ? w0 : ((i >> 2) == 1 ? w1 : ((i >> 2) == 2 ? w2 : w3));
          bool event = false;
          if ((uint32_t)i < nvalid) {
            state = delta[state * K + s_class[(word >> (8 * (i & 3))) & 0xFFu]];
            event = state >= thr && (uint32_t)i >= first_reported;
          }
          const uint32_t evmask = __ballot_sync (kFull, event);
          if (evmask)
            enqueue (evmask, event, i);
        }
      }
      pos += nvalid;
    }
    if (queued)
      drain (queued);
    __syncwarp ();
  }
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* Filter engine                                                                                                       */
/* ------------------------------------------------------------------------------------------------------------------ */
struct FilterParams {
  const void *text;
  uint64_t n, lead, base;
  uint32_t q;
  uint32_t tile_syms;  /* symbols per tile */
  uint64_t ntiles;
  const uint32_t *bloom;
  uint32_t bloom_words;
  const uint32_t *bloom2; /* optional second level in global memory */
  uint32_t bloom2_words;
  const acm_slot *qgrams;
  uint64_t qgram_mask;
  const uint4 *qset; /* widths 1 and 2: the q-gram keys as a compact set, 4 keys per 16-byte bucket (acm_tables.h) */
  uint32_t qset_shift, qset_has_empty_key;
  const acm_slot *edges;
  uint64_t edge_mask;
  const uint32_t *kw_len; /* tails: keyword id -> length, first symbol in the pool, the pool (forward symbols) */
  const uint64_t *kw_off;
  const void *kw_pool;
  const uint2 *kw_meta;     /* width 1 (else null): keyword id -> {length, first word of its reversed bytes in kw_rpool} */
  const uint32_t *kw_rpool;
  const uint32_t *prefix; /* symbols virtually preceding the text (carried cursor), prefix_len of them */
  uint32_t prefix_len;
  uint32_t stage_cap; /* raw hits a warp can stage per tile */
  /* F1 out */
  uint64_t *cand_pos;
  uint64_t cand_cap;
  unsigned long long *cand_count;
  uint64_t *tile_first;
  uint32_t *tile_n;
  uint32_t *overflow;
  /* F2 out / F4 in */
  uint32_t *cand_matches;
  uint32_t *cand_prefix; /* matches of the candidates before this one in its tile (written by F3) */
  uint4 *cand_inline;    /* the first two matches of the candidate {keyword, length, keyword, length}, shortest first (written by F2) */
  uint32_t *tile_matches;
  const uint64_t *tile_offsets;
  ACMB200Match *matches;
  uint64_t capacity;
  /* stride-2 kernel (F1s): its filter, its second level, and the spans it works in.  With this kernel the "tiles" of F2..F4
   * (tile_syms, ntiles, tile_first, tile_n) are the spans. */
  const uint32_t *bloom_s2;
  uint32_t bloom_s2_words;
  const uint32_t *s2_dist;   /* distance table (acm_tables.h), 1 << s2_dist_log2 words */
  uint32_t s2_dist_log2;
  const uint16_t *kw_dist;   /* keyword id -> chosen distances dA | dB << 8 (null: candidates carry no distance mask) */
  uint32_t lmax;             /* longest keyword */
  uint32_t *tile_spill;      /* F1s/F1h out (else null): how many of a span's candidates, the last ones, END in the next span */
  uint32_t s2_hit_cap;
  uint32_t *hot_spans;      /* spans F1s could not finish (a stage overflowed): redone exactly by filter_hot_spans_kernel */
  uint32_t hot_cap;
  unsigned int *hot_count;
  unsigned long long *span_counter;
};

/* The filter of a CTA, global -> shared memory, 16 bytes per load (both sides are 16-byte aligned; the last words one by one):
 * word by word the 175 KB took 43 rounds of dependent loads per thread, a fixed ~10 us of every scan. */
__device__ __forceinline__ void
copy_words_to_shared (uint32_t *dst, const uint32_t *__restrict__ src, uint32_t nwords) {
  const uint32_t nvec = nwords / 4;
  const uint4 *src4 = reinterpret_cast<const uint4 *> (src);
  uint4 *dst4 = reinterpret_cast<uint4 *> (dst);
#pragma unroll 4
  for (uint32_t i = threadIdx.x; i < nvec; i += blockDim.x)
    dst4[i] = __ldg (src4 + i);
  for (uint32_t i = nvec * 4 + threadIdx.x; i < nwords; i += blockDim.x)
    dst[i] = src[i];
}

template <int W> struct SymT;
template <> struct SymT<1> { typedef uint8_t type; };
template <> struct SymT<2> { typedef uint16_t type; };
template <> struct SymT<4> { typedef uint32_t type; };

/* symbol at (possibly negative) position: the carried-cursor prefix sits virtually before the text */
template <int W>
__device__ __forceinline__ bool
symbol_at (const FilterParams &p, int64_t pos, uint32_t *sym) {
  if (pos >= 0) {
    *sym = reinterpret_cast<const typename SymT<W>::type *> (p.text)[pos];
    return true;
  }
  if (pos >= -(int64_t)p.prefix_len) {
    *sym = p.prefix[(int64_t)p.prefix_len + pos];
    return true;
  }
  return false;
}

/* packed key of the q symbols ending at pos (last symbol most significant), false if they do not all exist */
template <int W>
__device__ __forceinline__ bool
qgram_key_at (const FilterParams &p, int64_t pos, uint64_t *key) {
  uint64_t k = 0;
  for (uint32_t j = 0; j < p.q; j++) {
    uint32_t s;
    if (!symbol_at<W> (p, pos - j, &s))
      return false;
    k = W == 4 ? ((j == 0 ? 0 : k << 32) | s) : ((k << (8 * W)) | s);
  }
  *key = k;
  return true;
}

__device__ __forceinline__ bool
slot_lookup (const acm_slot *__restrict__ tab, uint64_t mask, uint64_t key, uint32_t *node, uint32_t *keyword) {
  uint64_t j = acm_mix64 (key) & mask;
  for (;;) {
    const uint4 raw = __ldg (reinterpret_cast<const uint4 *> (tab + j));
    if (raw.z == ACM_TAB_NONE)
      return false;
    if ((((uint64_t)raw.y << 32) | raw.x) == key) {
      *node = raw.z;
      *keyword = raw.w;
      return true;
    }
    j = (j + 1) & mask;
  }
}

/* Filter test of the 16/W symbols a lane owns in one row.  Returns a mask: bit i = symbol i passed.
 * w[0] = the 4 bytes before the lane's 16, w[1..4] = the lane's 16 bytes. */
template <int W, int Q, int K, bool kTwoLevel>
__device__ __forceinline__ uint32_t
filter_row (const uint32_t *s_bloom, uint32_t nwords, const uint32_t *__restrict__ bloom2, uint32_t nwords2, const uint32_t (&w)[5]) {
  uint32_t acc = 0; /* every test shifts its verdict in at bit 31 (one funnel shift); the first symbol ends at bit 32 - 16/W */
  auto push = [&] (uint32_t folded) {
    const unsigned long long p1 = (unsigned long long)folded * ACM_BLOOM_C1;
    const uint32_t lo = (uint32_t)p1, hi = (uint32_t)(p1 >> 32);
    const uint32_t word = s_bloom[__umulhi (lo, nwords)];
    uint32_t t = (word >> (hi & 31u)) & (word >> (lo & 31u));
    if (kTwoLevel) { /* survivors of the shared-memory level ask the L2-resident level */
      uint32_t word2 = 0;
      if (t & 1u)
        word2 = __ldcg (bloom2 + __umulhi (folded * ACM_BLOOM_C3, nwords2)); /* L2 only: random table words must not evict text from L1 */
      const uint32_t h3 = __umulhi (folded, ACM_BLOOM_C4);
      t &= (word2 >> (h3 & 31u)) & (word2 >> ((h3 >> 5) & 31u));
    }
    acc = __funnelshift_r (acc, t, 1); /* (acc >> 1) | (t << 31): only bit 0 of t survives */
  };
  if (W == 1) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      /* the 4 bytes ending at byte i: b[i-3..i], b[i] most significant */
      const int j = (i >> 2) + 1, sh = ((i & 3) + 1) * 8;
      const uint32_t win = sh == 32 ? w[j] : __funnelshift_r (w[j - 1], w[j], sh);
      push (Q == 4 ? win : win >> (8 * (4 - Q)));
    }
  } else if (W == 2) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int j = (i >> 1) + 1;
      const uint32_t pair = (i & 1) ? w[j] : __funnelshift_r (w[j - 1], w[j], 16); /* low half = s[i-1], high half = s[i] */
      push (Q == 2 ? pair : (pair >> 16));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++)
      push (acm_fold_key (Q == 2 ? (((uint64_t)w[i + 1] << 32) | w[i]) : (uint64_t)w[i + 1]));
  }
  return acc >> (32 - 16 / W);
}

/* Membership of a 32-bit q-gram key in the compact set: normally one 16-byte load. */
__device__ __forceinline__ bool
qset_contains (const FilterParams &p, uint32_t key) {
  if (key == ACM_QSET_EMPTY)
    return p.qset_has_empty_key != 0;
  const uint32_t mask = (1u << (32 - p.qset_shift)) - 1u;
  for (uint32_t b = acm_qset_bucket (key, p.qset_shift);; b = (b + 1) & mask) {
    const uint4 c = __ldcg (p.qset + b); /* L2 only, as above */
    if (c.x == key || c.y == key || c.z == key || c.w == key)
      return true;
    if (c.x == ACM_QSET_EMPTY || c.y == ACM_QSET_EMPTY || c.z == ACM_QSET_EMPTY || c.w == ACM_QSET_EMPTY)
      return false;
  }
}

/* F1.  One warp per tile of kRows x 512 bytes; lane l of row r owns the 16 bytes at r*512 + l*16.
 * kOrdered: raw hits are staged in position order with a warp scan (used by the dense fallback, where a tile may hold thousands
 * of candidates); otherwise they are staged through a shared-memory counter in any order and the few survivors of the exact
 * confirmation are sorted afterwards -- fewer instructions when hits are rare. */
template <int W, int kRows, int Q, int K, bool kOrdered, bool kTwoLevel>
__global__ void __launch_bounds__ (1024, 1)
filter_scan_kernel (const __grid_constant__ FilterParams p) {
  extern __shared__ __align__ (16) unsigned char smem[];
  uint32_t *s_bloom = reinterpret_cast<uint32_t *> (smem);
  unsigned char *s_stage_all = smem + (size_t)p.bloom_words * 4;
  copy_words_to_shared (s_bloom, p.bloom, p.bloom_words);
  __syncthreads ();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const uint32_t lanes_below = (1u << lane) - 1u;
  constexpr int kSyms = 16 / W;        /* symbols per lane per row */
  constexpr int kRowSyms = 32 * kSyms; /* symbols per row */
  constexpr uint32_t kTileSyms = kRows * kRowSyms; /* symbols per tile: kRows loads in flight per lane */
  /* per warp: a 16-byte header (hit counter) then stage_cap tile-relative positions -- 16 bits each, 8 in the dense mode when a
   * tile has at most 256 symbols (its stage holds EVERY position of a tile: with bytes, two rows of 32-bit symbols take the room
   * of one, and the fixed costs of a tile are spread over twice the text) */
  typedef typename std::conditional<(kOrdered && kTileSyms <= 256), uint8_t, uint16_t>::type StageT;
  uint32_t *stage_count = reinterpret_cast<uint32_t *> (s_stage_all + (size_t)warp * (p.stage_cap * sizeof (StageT) + 16));
  StageT *stage = reinterpret_cast<StageT *> (stage_count + 4);
  constexpr uint32_t q = Q;
  const uint32_t nwords = p.bloom_words, stage_cap = p.stage_cap;
  const uint64_t first_valid = max (p.lead, (uint64_t)(q - 1)); /* windows that start before the text are handled below */
  const uint8_t *text8 = reinterpret_cast<const uint8_t *> (p.text);

  /* The rows of a tile live in registers; the loads of the NEXT tile are issued right after the current tile has been filtered
   * and staged (its rows are dead by then), so their latency hides behind the confirmation step instead of stalling the filter. */
  uint4 v[kRows];
  uint32_t before_tile = 0;
  const uint64_t tile_stride = (uint64_t)gridDim.x * warps;
  /* interior also means: the aligned word after the tile is inside the text (the confirmation step reads it for the tile's last symbols) */
  auto interior_tile = [&] (uint64_t t) { return t * kTileSyms >= first_valid + 4 && (t + 1) * kTileSyms + 4 <= p.n; };
  auto load_tile = [&] (uint64_t t) {
    const uint64_t base = t * kTileSyms;
    const uint8_t *ptr = text8 + base * W;
    if (interior_tile (t)) {
#pragma unroll
      for (int r = 0; r < kRows; r++)
        v[r] = *reinterpret_cast<const uint4 *> (ptr + r * 512 + lane * 16);
    } else {
#pragma unroll
      for (int r = 0; r < kRows; r++) {
        const uint64_t byte0 = (base + (uint64_t)r * kRowSyms + (uint64_t)lane * kSyms) * W, nbytes = p.n * W;
        if (byte0 + 16 <= nbytes)
          v[r] = *reinterpret_cast<const uint4 *> (text8 + byte0);
        else {
          uint32_t w[4] = { 0, 0, 0, 0 };
          for (int i = 0; i < 16; i++)
            if (byte0 + i < nbytes)
              w[i >> 2] |= (uint32_t)text8[byte0 + i] << (8 * (i & 3));
          v[r] = make_uint4 (w[0], w[1], w[2], w[3]);
        }
      }
    }
    before_tile = base * W >= 4 ? *reinterpret_cast<const uint32_t *> (ptr - 4) : 0;
  };
  {
    const uint64_t first_tile = (uint64_t)blockIdx.x * warps + warp;
    if (first_tile < p.ntiles)
      load_tile (first_tile);
  }

  for (uint64_t tile = (uint64_t)blockIdx.x * warps + warp; tile < p.ntiles; tile += tile_stride) {
    const uint64_t tile_base = tile * kTileSyms; /* in symbols */
    uint32_t staged = 0; /* warp-uniform */
    { /* pull the tile after the next one from DRAM into L2 */
      const uint64_t ahead = tile + 2 * tile_stride;
      if (ahead < p.ntiles && lane < kRows * 4)
        asm volatile ("prefetch.global.L2 [%0];" ::"l"(text8 + ahead * (kTileSyms * W) + lane * 128));
    }
    /* interior tiles (every symbol reportable, every vector load inside the text) take the check-free path */
    const bool interior = interior_tile (tile);

    /* positions whose window reaches into the carried-cursor prefix: checked exactly, by lane 0 of the first tile */
    if (tile == 0 && p.prefix_len && q > 1) {
      if (lane == 0)
        for (uint64_t pos = p.lead; pos < min ((uint64_t)(q - 1), p.n); pos++) {
          uint64_t key;
          if (qgram_key_at<W> (p, (int64_t)pos, &key) && staged < stage_cap)
            stage[staged++] = (StageT)pos; /* confirmed below like any other staged hit */
        }
      staged = __shfl_sync (kFull, staged, 0);
    }
    if (!kOrdered) {
      if (lane == 0)
        *stage_count = staged;
      __syncwarp ();
    }

    uint32_t hits[kRows];
#pragma unroll
    for (int r = 0; r < kRows; r++) {
      const uint32_t up = __shfl_up_sync (kFull, v[r].w, 1);
      const uint32_t wrap = r == 0 ? before_tile : __shfl_sync (kFull, v[r > 0 ? r - 1 : 0].w, 31);
      const uint32_t w[5] = { lane == 0 ? wrap : up, v[r].x, v[r].y, v[r].z, v[r].w };
      hits[r] = filter_row<W, Q, K, kTwoLevel> (s_bloom, nwords, p.bloom2, p.bloom2_words, w);
    }
    if (!interior) { /* drop positions outside [first_valid, n) */
#pragma unroll
      for (int r = 0; r < kRows; r++) {
        const uint64_t pos0 = tile_base + (uint64_t)r * kRowSyms + (uint64_t)lane * kSyms;
        uint32_t keep = 0;
        for (int i = 0; i < kSyms; i++)
          if (pos0 + i >= first_valid && pos0 + i < p.n)
            keep |= 1u << i;
        hits[r] &= keep;
      }
    }

    if (kOrdered) {
      /* Ordered append.  Position order is (row, lane, symbol): one warp scan over the per-row counts of every lane, two rows
       * packed per 32-bit word (a row holds at most 512 hits). */
      uint32_t incl[(kRows + 1) / 2];
#pragma unroll
      for (int h = 0; h < (kRows + 1) / 2; h++)
        incl[h] = __popc (hits[2 * h]) | ((2 * h + 1 < kRows ? __popc (hits[2 * h + 1]) : 0) << 16);
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int h = 0; h < (kRows + 1) / 2; h++) {
          const uint32_t o = __shfl_up_sync (kFull, incl[h], d);
          if (lane >= d)
            incl[h] += o;
        }
      }
      uint32_t row_start = staged;
#pragma unroll
      for (int r = 0; r < kRows; r++) {
        const uint32_t mine = (incl[r >> 1] >> (16 * (r & 1))) & 0xFFFFu; /* inclusive count up to this lane */
        const uint32_t total = __shfl_sync (kFull, mine, 31);
        uint32_t at = row_start + mine - __popc (hits[r]);
        uint32_t h = hits[r];
        const uint32_t rel0 = (uint32_t)(r * kRowSyms + lane * kSyms);
        while (h) {
          const int i = __ffs (h) - 1;
          h &= h - 1;
          if (at < stage_cap)
            stage[at] = (StageT)(rel0 + i);
          at++;
        }
        row_start += total;
      }
      staged = row_start;
      __syncwarp ();
    } else {
      /* Unordered append through the warp's shared counter: only lanes that have hits do any work. */
      uint32_t mine = 0;
#pragma unroll
      for (int r = 0; r < kRows; r++)
        mine += __popc (hits[r]);
      if (mine) {
        uint32_t at = atomicAdd (stage_count, mine);
#pragma unroll
        for (int r = 0; r < kRows; r++) {
          uint32_t h = hits[r];
          const uint32_t rel0 = (uint32_t)(r * kRowSyms + lane * kSyms);
          while (h) {
            const int i = __ffs (h) - 1;
            h &= h - 1;
            if (at < stage_cap)
              stage[at] = (StageT)(rel0 + i);
            at++;
          }
        }
      }
    }
    if (tile + tile_stride < p.ntiles)
      load_tile (tile + tile_stride); /* in flight during the confirmation below */
    if (!kOrdered) {
      __syncwarp ();
      staged = *stage_count;
    }
    if (staged > stage_cap) {
      if (lane == 0)
        atomicExch (p.overflow, 1u);
      staged = stage_cap;
    }

    /* Exact confirmation of the staged hits in the q-gram table; survivors are compacted in place (stable). Two batches are
     * in flight at a time so that the dependent L2 round trips (text, then table) overlap. */
    uint32_t kept = 0;
    for (uint32_t b = 0; b < staged; b += 64) {
      uint32_t rel[2] = { 0, 0 }, key[2] = { 0, 0 };
      bool live[2], ok[2] = { false, false };
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const uint32_t i = b + 32 * u + lane;
        live[u] = i < staged;
        if (live[u]) {
          rel[u] = stage[i];
          if (W != 4 && interior) { /* the Q symbols ending at the position, rebuilt from two aligned 32-bit loads inside the tile */
            const uint64_t first4 = (tile_base + rel[u]) * W + (W - 1) - 3;
            const uint32_t *t32 = reinterpret_cast<const uint32_t *> (text8 + (first4 & ~(uint64_t)3));
            const uint32_t win = __funnelshift_r (t32[0], t32[1], 8 * (uint32_t)(first4 & 3));
            key[u] = W == 1 ? (Q == 4 ? win : win >> (8 * (4 - Q))) : (Q == 2 ? win : win >> 16);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; u++)
        if (live[u]) {
          if (W != 4 && interior)
            ok[u] = qset_contains (p, key[u]);
          else {
            uint64_t k64;
            uint32_t node, kw;
            ok[u] = qgram_key_at<W> (p, (int64_t)(tile_base + rel[u]), &k64) && slot_lookup (p.qgrams, p.qgram_mask, k64, &node, &kw);
          }
        }
      __syncwarp ();
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const uint32_t mask = __ballot_sync (kFull, ok[u]);
        if (ok[u])
          stage[kept + __popc (mask & lanes_below)] = (StageT)rel[u];
        kept += __popc (mask);
      }
      __syncwarp ();
    }
    if (!kOrdered && kept > stage_cap / 2) { /* the sort below needs half of the stage as scratch: leave it to the dense mode */
      if (lane == 0)
        atomicExch (p.overflow, 1u);
      kept = 0;
    }
    if (!kOrdered && kept > 1) {
      /* sort the survivors by position (they are few): rank = number of smaller positions */
      for (uint32_t b = 0; b < kept; b += 32) {
        const uint32_t i = b + lane;
        uint32_t mine = i < kept ? stage[i] : 0xFFFFFFFFu, rank = 0;
        for (uint32_t j = 0; j < kept; j++)
          rank += stage[j] < mine;
        __syncwarp ();
        /* positions are distinct, so ranks are a permutation; write through a second pass to avoid overwriting unread entries */
        if (b + 32 >= kept && b == 0) {
          if (i < kept)
            stage[rank] = (StageT)mine;
        } else { /* more than 32 survivors: use the upper half of the stage as scratch */
          if (i < kept)
            stage[stage_cap / 2 + rank] = (StageT)mine;
        }
        __syncwarp ();
      }
      if (kept > 32) {
        for (uint32_t i = lane; i < kept; i += 32)
          stage[i] = stage[stage_cap / 2 + i];
        __syncwarp ();
      }
    }
    /* one reservation per tile; the tile's candidates stay contiguous and ordered */
    uint64_t first = 0;
    if (kept) {
      unsigned long long seg = 0;
      if (lane == 0)
        seg = atomicAdd (p.cand_count, (unsigned long long)kept);
      seg = __shfl_sync (kFull, seg, 0);
      if (seg + kept > p.cand_cap) {
        if (lane == 0)
          atomicExch (p.overflow, 1u);
        kept = 0;
      } else {
        first = seg;
        for (uint32_t i = lane; i < kept; i += 32)
          p.cand_pos[seg + i] = tile_base + stage[i];
      }
    }
    if (lane == 0) {
      p.tile_first[tile] = first;
      p.tile_n[tile] = kept;
    }
    __syncwarp ();
  }
}

/* F1s: the stride-2 variant of F1 for byte alphabets whose shortest keyword has at least 4 bytes (acm_tables.h).
 *
 * Only the even offsets of a tile are tested, each on the 3-byte window ending there: 8 tests per lane and 512-byte row instead
 * of 16, built with one byte permute each (one shared-memory lookup per TWO text bytes).  The filter holds, for every keyword, the
 * two windows the finalise step chose for it; a hit at the sampled position s is confirmed in the distance table (L2, one word
 * load, ld.global.cg): an entry that matches the window's extension byte says that a keyword may END at s + d.  Hits are staged as
 * (lane, test index), slots from a warp scan; the confirmation re-reads the 5 bytes s-3..s+1 (usually still in L1: the tile
 * loads are evict_last and the kernel leaves ~60 KB of the SM to L1).  A warp works through a span of consecutive tiles taken
 * from a global counter, sorts the span's candidates by end position (merging the distance masks of equal ends) and appends them
 * with one reservation -- so F2..F4 see spans where they saw tiles; the candidates whose end lies in the NEXT span come last and
 * are counted in tile_spill, F3 merges them into the next span's order.  F2 verifies candidates exactly, so everything here may
 * err on the side of keeping a position; nothing may drop one. */
constexpr uint32_t kS2SpanBytes = 32768;
constexpr uint32_t kS2EndSlack = 32; /* a candidate end lies at most ACM_S2_DMAX bytes after its sampled position */

/* Exact (slow) form of the confirmation, for the first / last tile and for the spans redone by filter_hot_spans_kernel:
 * calls emit (d) for every distance-table entry that matches the window ending at the sampled position s (2 <= s < n). */
template <typename F>
__device__ __forceinline__ void
s2_probe (const FilterParams &p, const uint8_t *text8, uint64_t s, F &&emit) {
  const uint32_t gram3 = (uint32_t)text8[s - 2] | ((uint32_t)text8[s - 1] << 8) | ((uint32_t)text8[s] << 16);
  const int left = s >= 3 ? (int)text8[s - 3] : -1, right = s + 1 < p.n ? (int)text8[s + 1] : -1;
  const uint32_t mask = (1u << p.s2_dist_log2) - 1u;
  for (uint32_t idx = acm_pair_word (gram3, p.s2_dist_log2);; idx = (idx + 1) & mask) {
    const uint32_t word = __ldcg (p.s2_dist + idx);
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const uint32_t ent = (word >> (16 * h)) & 0xFFFFu;
      if ((ent & ACM_S2D_VALID) && ((ent & ACM_S2D_RIGHT) ? right : left) == (int)(ent & 0xFFu))
        emit (ACM_S2D_DIST (ent));
    }
    if (!(word & ACM_S2D_CONT))
      break;
  }
}

/* End of a span of the stride-2 kernels: the span's candidates (cands[], any order) are sorted by end position, equal ends merged,
 * and appended to the candidate list with one reservation; a span whose stages overflowed (hot) is listed for
 * filter_hot_spans_kernel instead. */
__device__ __forceinline__ void
s2_finish_span (const FilterParams &p, uint32_t span, uint64_t span_base, uint32_t *cands, uint32_t ncand, bool hot, int lane) {
  /* ---- the span's candidates: sorted by end position, equal ends merged (their distance masks OR-ed) ---- */
  uint32_t nspill = 0;
  if (hot) { /* rare */
    ncand = 0;
    if (lane == 0) {
      const unsigned int at = atomicAdd (p.hot_count, 1u);
      if (at < p.hot_cap)
        p.hot_spans[at] = (uint32_t)span;
    }
  } else if (ncand) {
    uint32_t val[2], dm[2], rank[2] = { 0, 0 };
    bool first[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const uint32_t i = lane + 32 * r;
      first[r] = i < ncand;
      val[r] = first[r] ? cands[i] : 0xFFFFFFFFu;
      dm[r] = val[r] & ACM_S2_DMASK_ALL;
    }
    if (ncand > 1) {
      for (uint32_t j = 0; j < ncand; j++) {
        const uint32_t w = cands[j];
#pragma unroll
        for (int r = 0; r < 2; r++)
          if ((w >> 15) == (val[r] >> 15)) {
            dm[r] |= w & ACM_S2_DMASK_ALL;
            if (j < (uint32_t)lane + 32 * r)
              first[r] = false; /* an earlier entry with the same end represents it */
          }
      }
      const uint32_t f0 = __ballot_sync (kFull, first[0]), f1 = __ballot_sync (kFull, first[1]);
      for (uint32_t j = 0; j < ncand; j++)
        if (((j < 32 ? f0 >> j : f1 >> (j - 32)) & 1u)) {
          const uint32_t w = cands[j];
#pragma unroll
          for (int r = 0; r < 2; r++)
            rank[r] += (w >> 15) < (val[r] >> 15);
        }
      __syncwarp ();
#pragma unroll
      for (int r = 0; r < 2; r++)
        if (first[r])
          cands[rank[r]] = (val[r] & ~ACM_S2_DMASK_ALL) | dm[r];
      ncand = __popc (f0) + __popc (f1);
      __syncwarp ();
    }
    /* ends beyond the span (at most kS2EndSlack bytes into the next one) come last */
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const uint32_t i = lane + 32 * r;
      nspill += __popc (__ballot_sync (kFull, i < ncand && (cands[i] >> 15) >= kS2SpanBytes));
    }
  }

  /* ---- one reservation per span ---- */
  uint64_t first_slot = 0;
  if (ncand) {
    unsigned long long seg = 0;
    if (lane == 0)
      seg = atomicAdd (p.cand_count, (unsigned long long)ncand);
    seg = __shfl_sync (kFull, seg, 0);
    if (seg + ncand > p.cand_cap) {
      if (lane == 0)
        atomicExch (p.overflow, 1u);
      ncand = nspill = 0;
    } else {
      first_slot = seg;
      for (uint32_t i = lane; i < ncand; i += 32) {
        const uint32_t c = cands[i];
        p.cand_pos[seg + i] = (span_base + (c >> 15)) | ((uint64_t)(c & ACM_S2_DMASK_ALL) << 48);
      }
    }
  }
  if (lane == 0) {
    p.tile_first[span] = first_slot;
    p.tile_n[span] = ncand;
    p.tile_spill[span] = nspill;
  }
  __syncwarp ();
}

template <int K, int kBatches, int kRows>
__global__ void __launch_bounds__ (1024, 1)
filter_scan_s2_kernel (const __grid_constant__ FilterParams p) {
  constexpr uint32_t kTileBytes = kRows * 512;
  constexpr int kAcc = (kRows + 3) / 4;       /* 32 verdict bits per accumulator: 8 tests per row */
  constexpr int kAccShift = kRows < 4 ? 32 - 8 * kRows : 0; /* fewer than 32 tests: move the verdicts to the top bits */
  constexpr uint32_t kSpanTiles = kS2SpanBytes / kTileBytes;
  extern __shared__ __align__ (16) unsigned char smem[];
  uint32_t *s_bloom = reinterpret_cast<uint32_t *> (smem);
  copy_words_to_shared (s_bloom, p.bloom_s2, p.bloom_s2_words);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lanes_below = (1u << lane) - 1u;
  const uint32_t kHitCap = p.s2_hit_cap; /* 1.5x the expected hits per tile + 32, see acm_finalise.c */
  unsigned char *mine = smem + (size_t)p.bloom_s2_words * 4 + (size_t)warp * ACM_S2_WARP_BYTES (kHitCap);
  uint32_t *slow_count = reinterpret_cast<uint32_t *> (mine); /* candidate counter of the slow (first / last tile) path */
  uint16_t *hits = reinterpret_cast<uint16_t *> (mine + 16);
  uint32_t *cands = reinterpret_cast<uint32_t *> (mine + 16 + kHitCap * 2); /* (span-relative end << 15) | distance mask */
  __syncthreads ();

  const uint32_t nwords = p.bloom_s2_words;
  const uint8_t *text8 = reinterpret_cast<const uint8_t *> (p.text);
  const uint32_t dist_shift = 32 - p.s2_dist_log2, dist_mask = (1u << p.s2_dist_log2) - 1u;
  const uint64_t ntiles = (p.n + kTileBytes - 1) / kTileBytes;
  uint4 v[kRows];
  uint32_t before_tile = 0;
  /* interior: every candidate end of the tile is reportable and inside the text, every window is preceded by text */
  auto interior_tile = [&] (uint64_t t) { return t * kTileBytes >= max (p.lead, (uint64_t)4) && (t + 1) * kTileBytes + kS2EndSlack <= p.n; };
  /* chained: v[] still holds tile t - 1, whose last word (lane 31's) is the word before tile t: no load for it */
  auto load_tile = [&] (uint64_t t, bool chained) {
    const uint64_t base = t * kTileBytes;
    const uint8_t *ptr = text8 + base;
    if (base >= 4 && (t + 1) * kTileBytes <= p.n) {
      before_tile = chained ? v[kRows - 1].w : *reinterpret_cast<const uint32_t *> (ptr - 4); /* only lane 31's copy is used */
#pragma unroll
      for (int r = 0; r < kRows; r++) /* evict_last: the confirmation step re-reads a few words of the tile one iteration later */
        asm volatile ("ld.global.L1::evict_last.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[r].x), "=r"(v[r].y), "=r"(v[r].z), "=r"(v[r].w) : "l"(ptr + r * 512 + lane * 16));
    } else { /* first / last tile: bytes outside the text read as zero */
#pragma unroll
      for (int r = 0; r < kRows; r++) {
        const uint64_t byte0 = base + (uint64_t)r * 512 + (uint64_t)lane * 16;
        if (byte0 + 16 <= p.n)
          v[r] = *reinterpret_cast<const uint4 *> (text8 + byte0);
        else {
          uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0; /* not unrolled: this path runs for one or two tiles per scan */
#pragma unroll 1
          for (int i = 15; i >= 0; i--) {
            const uint32_t b = byte0 + i < p.n ? text8[byte0 + i] : 0u;
            w3 = (w3 << 8) | (w2 >> 24);
            w2 = (w2 << 8) | (w1 >> 24);
            w1 = (w1 << 8) | (w0 >> 24);
            w0 = (w0 << 8) | b;
          }
          v[r] = make_uint4 (w0, w1, w2, w3);
        }
      }
      before_tile = base >= 4 ? *reinterpret_cast<const uint32_t *> (ptr - 4) : 0;
    }
  };

  for (;;) {
    /* the next span, from the global counter; broadcast with a warp reduction (one instruction; its result is warp-uniform as
     * far as the compiler is concerned).  The host keeps the number of spans below 2^32. */
    uint32_t mine_span = 0;
    if (lane == 0) {
      const unsigned long long got = atomicAdd (p.span_counter, 1ull);
      mine_span = got >= p.ntiles ? 0xFFFFFFFFu : (uint32_t)got;
    }
    const uint32_t span = __reduce_max_sync (kFull, mine_span);
    if (span == 0xFFFFFFFFu)
      break;
    const uint64_t tile0 = (uint64_t)span * kSpanTiles, tile1 = min (ntiles, tile0 + kSpanTiles);
    uint32_t ncand = 0; /* warp-uniform: candidates of this span held in cands[], in no particular order until the span is done */
    bool hot = false;   /* warp-uniform: a stage of this span overflowed, the span is left to filter_hot_spans_kernel */
    load_tile (tile0, false);

    /* occurrences that begin in the carried-cursor prefix end within the first lmax - 1 positions: those ends are candidates
     * for every distance (F2 decides) */
    if (span == 0 && p.prefix_len) {
      const uint64_t last = min ((uint64_t)(p.lmax ? p.lmax - 1 : 0), p.n);
      for (uint64_t e0 = p.lead; e0 < last; e0 += 32) {
        const uint64_t e = e0 + lane;
        const bool keep = e < last;
        const uint32_t m = __ballot_sync (kFull, keep);
        const uint32_t at = ncand + __popc (m & lanes_below);
        if (keep && at < ACM_S2_CAND_CAP)
          cands[at] = ((uint32_t)e << 15) | ACM_S2_DMASK_ALL;
        ncand += __popc (m);
      }
      __syncwarp ();
    }

    hot = ncand > ACM_S2_CAND_CAP;
    for (uint64_t tile = tile0; tile < tile1 && !hot; tile++) {
      const uint64_t tile_base = tile * kTileBytes;
      const bool interior = interior_tile (tile);
      const uint32_t tile_off = (uint32_t)(tile - tile0) * kTileBytes;
      if (tile + 2 < tile1 && lane < kRows * 4)
        asm volatile ("prefetch.global.L2 [%0];" ::"l"(text8 + (tile + 2) * kTileBytes + lane * 128));

      /* ---- filter: 8 tests per lane and row, verdicts shifted into acc[] (test 0 of an accumulator ends at bit 31) ---- */
      uint32_t acc[kAcc];
#pragma unroll
      for (int a = 0; a < kAcc; a++)
        acc[a] = 0;
      auto test = [&] (uint32_t key, uint32_t &dst) {
        const unsigned long long p1 = (unsigned long long)key * ACM_BLOOM_C1;
        const uint32_t lo = (uint32_t)p1, hi = (uint32_t)(p1 >> 32);
        const uint32_t word = s_bloom[__umulhi (lo, nwords)];
        const uint32_t t = (word >> (hi & 31u)) & (word >> (lo & 31u)) & 1u;
        asm ("mad.lo.u32 %0, %0, 2, %1;" : "+r"(dst) : "r"(t)); /* dst = 2 dst + verdict on the FMA pipe (kept out of the ALU-side LOP3 trees) */
      };
#pragma unroll
      for (int r = 0; r < kRows; r++) {
        /* the word before the lane's 16 bytes comes from the lane below; lane 0 needs the last word of the row above, which
         * lane 31 sends in place of its own: ONE rotating shuffle per row */
        const uint32_t send = lane == 31 ? (r == 0 ? before_tile : v[r > 0 ? r - 1 : 0].w) : v[r].w;
        const uint32_t w[5] = { __shfl_sync (kFull, send, (lane + 31) & 31), v[r].x, v[r].y, v[r].z, v[r].w };
#pragma unroll
        for (int j = 0; j < 4; j++) {
          test (__byte_perm (w[j], w[j + 1], 0x4432), acc[r >> 2]); /* window ending at byte 4j of the lane's 16: two bytes of the word before */
          test (__byte_perm (w[j + 1], 0u, 0x2210), acc[r >> 2]);   /* window ending at byte 4j+2 */
        }
      }

      if (!interior) { /* first / last tile: drop the tests whose window is not wholly inside the text (zero padding: a thousand equal keys) */
#pragma unroll
        for (int a = 0; a < kAcc; a++) {
          uint32_t keep = 0;
          for (int i = 0; i < 32 && a * 32 + i < 8 * kRows; i++) {
            const uint32_t ti = a * 32 + i;
            const uint64_t s = tile_base + (uint64_t)lane * 16 + (uint64_t)(ti >> 3) * 512 + (ti & 7u) * 2;
            if (s >= 2 && s < p.n)
              keep |= 0x80000000u >> i;
          }
          acc[a] &= keep >> kAccShift;
        }
      }

      /* ---- stage the hits as (lane << 6) | test index; slots from a warp scan of the per-lane counts (a shared counter would
       * serialise the lanes that have hits on one bank) ---- */
      uint32_t staged;
      {
        uint32_t cnt = 0;
#pragma unroll
        for (int a = 0; a < kAcc; a++)
          cnt += __popc (acc[a]);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t o = __shfl_up_sync (kFull, incl, d);
          if (lane >= d)
            incl += o;
        }
        staged = __shfl_sync (kFull, incl, 31);
        if (staged > kHitCap) { /* a run of equal bytes that happens to be a key, say: more hits than the stage holds */
          hot = true;
          staged = 0;
#pragma unroll
          for (int a = 0; a < kAcc; a++)
            acc[a] = 0;
        }
        uint32_t at = incl - cnt;
#pragma unroll
        for (int a = 0; a < kAcc; a++) {
          const uint32_t tag = ((uint32_t)lane << 6) | (a * 32);
          uint32_t bits = acc[a] << kAccShift;
          while (bits) {
            const uint32_t ti = __clz (bits);
            bits &= ~(0x80000000u >> ti);
            hits[at++] = (uint16_t)(tag | ti);
          }
        }
      }
      if (tile + 1 < tile1)
        load_tile (tile + 1, true); /* in flight during the confirmation below */
      __syncwarp ();

      /* appends the end s + d of every lane whose entry matched, in lane order */
      auto append = [&] (bool ok, uint32_t rel, uint32_t ent) {
        const uint32_t m = __ballot_sync (kFull, ok);
        if (m) { /* rare */
          const uint32_t at = ncand + __popc (m & lanes_below), d = ACM_S2D_DIST (ent);
          if (ok && at < ACM_S2_CAND_CAP)
            cands[at] = ((tile_off + rel + d) << 15) | (1u << (d >> 1));
          ncand += __popc (m);
        }
      };

      /* ---- confirmation: up to kBatches x 32 hits in flight.  Interior tiles run straight-line code: every lane loads (idle lanes
       * repeat the last hit) so that all text loads are issued before the first one is consumed, then all distance-table loads. ---- */
      auto confirm = [&] (auto nb_tag, uint32_t b) {
        constexpr int kNb = decltype (nb_tag)::value;
        uint32_t rel[kNb];
        bool live[kNb];
#pragma unroll
        for (int u = 0; u < kNb; u++) {
          const uint32_t i = b + 32 * u + lane;
          live[u] = i < staged;
          const uint32_t h = hits[min (i, staged - 1)], ti = h & 63u;
          rel[u] = ((h >> 6) << 4) + ((ti >> 3) << 9) + ((ti & 7u) << 1); /* the sampled position, tile-relative */
        }
        if (interior) {
          uint32_t t0[kNb], t1[kNb], word[kNb], key[kNb];
          const uint8_t *tile_m4 = text8 + tile_base - 4; /* the tile is preceded by text */
#pragma unroll
          for (int u = 0; u < kNb; u++) { /* bytes s-3 .. s from two aligned words */
            const uint32_t *t32 = reinterpret_cast<const uint32_t *> (tile_m4 + (uint64_t)((rel[u] + 1u) & ~3u));
            t0[u] = t32[0]; /* through L1: the tile's lines are often still there */
            t1[u] = t32[1];
          }
#pragma unroll
          for (int u = 0; u < kNb; u++) {
            key[u] = __funnelshift_r (t0[u], t1[u], 8u * ((rel[u] + 1u) & 3u)); /* (s - 3) mod 4 = 1 or 3 */
            word[u] = __ldcg (p.s2_dist + (((key[u] >> 8) * ACM_PAIR_C0) >> dist_shift)); /* L2 only: random words must not evict the tiles from L1 */
          }
#pragma unroll
          for (int u = 0; u < kNb; u++) {
            /* usual entries: valid, extended to the LEFT by the window's first byte -- both halves of the word in one xor/and
             * (VALID 0x8000 set, RIGHT 0x2000 clear, extension byte equal) */
            const uint32_t x = (word[u] ^ ((key[u] & 0xFFu) * 0x00010001u | 0x80008000u)) & 0xA0FFA0FFu;
            const bool ok0 = live[u] && (x & 0xFFFFu) == 0, ok1 = live[u] && (x >> 16) == 0;
            if (__any_sync (kFull, ok0 || ok1)) { /* about one candidate per two tiles */
              append (ok0, rel[u], word[u] & 0xFFFFu);
              append (ok1, rel[u], word[u] >> 16);
            }
            /* rare: an entry extended to the RIGHT (the first window of a 4-byte keyword: VALID is bit 15 and RIGHT bit 13 of an
             * entry) or a word that continues in the next one */
            const bool slow = live[u] && ((word[u] & (word[u] << 2) & 0x80008000u) | (word[u] & ACM_S2D_CONT)) != 0;
            if (__any_sync (kFull, slow)) {
              const uint32_t left = key[u] & 0xFFu;
              uint32_t right = 0x100u, idx = ((key[u] >> 8) * ACM_PAIR_C0) >> dist_shift, wd = word[u];
              bool more = slow, first = true; /* the left-extended entries of the first word are done */
              for (;;) {
                const uint32_t e0 = wd & 0xFFFFu, e1 = wd >> 16;
                const bool v0 = more && (e0 & ACM_S2D_VALID) && (!first || (e0 & ACM_S2D_RIGHT)), v1 = more && (e1 & ACM_S2D_VALID) && (!first || (e1 & ACM_S2D_RIGHT));
                if (((v0 && (e0 & ACM_S2D_RIGHT)) || (v1 && (e1 & ACM_S2D_RIGHT))) && right == 0x100u)
                  right = text8[tile_base + rel[u] + 1];
                append (v0 && (e0 & 0xFFu) == ((e0 & ACM_S2D_RIGHT) ? right : left), rel[u], e0);
                append (v1 && (e1 & 0xFFu) == ((e1 & ACM_S2D_RIGHT) ? right : left), rel[u], e1);
                more = more && (wd & ACM_S2D_CONT);
                first = false;
                if (!__any_sync (kFull, more))
                  break;
                idx = (idx + 1u) & dist_mask;
                wd = more ? __ldcg (p.s2_dist + idx) : 0u;
              }
            }
          }
        } else { /* first / last tile: every window and every end is checked for its range */
          if (lane == 0)
            *slow_count = ncand;
          __syncwarp ();
#pragma unroll
          for (int u = 0; u < kNb; u++) {
            const uint64_t s = tile_base + rel[u];
            if (live[u] && s >= 2 && s < p.n)
              s2_probe (p, text8, s, [&] (uint32_t d) {
                const uint64_t e = s + d;
                if (e >= p.lead && e < p.n) {
                  const uint32_t at = atomicAdd (slow_count, 1u);
                  if (at < ACM_S2_CAND_CAP)
                    cands[at] = ((tile_off + rel[u] + d) << 15) | (1u << (d >> 1));
                }
              });
          }
          __syncwarp ();
          ncand = *slow_count;
          __syncwarp ();
        }
      };
      {
        uint32_t b = 0; /* full groups of kBatches batches, then what is left as one group of 2 or 1 (kBatches <= 3) */
        static_assert (kBatches >= 1 && kBatches <= 3, "the remainder handling below covers at most two left-over batches");
        for (; b + 32 * (kBatches - 1) < staged; b += 32 * kBatches)
          confirm (std::integral_constant<int, kBatches> (), b);
        if (kBatches > 2 && b + 32 < staged) {
          confirm (std::integral_constant<int, 2> (), b);
          b += 64;
        }
        if (kBatches > 1 && b < staged)
          confirm (std::integral_constant<int, 1> (), b);
      }
      __syncwarp ();
      if (ncand > ACM_S2_CAND_CAP)
        hot = true;
    }

    s2_finish_span (p, span, tile0 * kTileBytes, cands, ncand, hot, lane);
  }
}

/* F1h: the spans F1s left unfinished (more filter hits in a tile, or more candidates in the span, than its shared-memory stages
 * hold).  One warp per such span probes the distance table for EVERY sampled position of the span (no filter), ORs the
 * distance masks of the ends into a span-sized array in shared memory, reserves the span's segment of the candidate list once and
 * writes the ends in order -- what F1s would have produced, without its capacity limits.  The grid is fixed; the number of
 * unfinished spans is read from device memory (no host round trip between F1s and this kernel). */
constexpr uint32_t kHotSmemBytes = (kS2SpanBytes + kS2EndSlack) * 2;
__global__ void __launch_bounds__ (32)
filter_hot_spans_kernel (const __grid_constant__ FilterParams p) {
  extern __shared__ __align__ (16) unsigned char smem[];
  uint32_t *s_mask32 = reinterpret_cast<uint32_t *> (smem); /* 16-bit distance mask per span-relative end, two per word */
  const int lane = threadIdx.x;
  const uint32_t nb_hot = min (*p.hot_count, p.hot_cap);
  const uint8_t *text8 = reinterpret_cast<const uint8_t *> (p.text);
  constexpr uint32_t kEnds = kS2SpanBytes + kS2EndSlack;
  for (uint32_t h = blockIdx.x; h < nb_hot; h += gridDim.x) {
    const uint64_t span = p.hot_spans[h], span_base = span * kS2SpanBytes;
    const uint64_t span_end = min (p.n, span_base + kS2SpanBytes);
    for (uint32_t i = lane; i < kEnds / 2; i += 32)
      s_mask32[i] = 0;
    __syncwarp ();
    auto mark = [&] (uint64_t e, uint32_t dmask) {
      const uint32_t rel = (uint32_t)(e - span_base);
      atomicOr (&s_mask32[rel >> 1], dmask << (16 * (rel & 1u)));
    };
    if (span == 0 && p.prefix_len) {
      const uint64_t last = min ((uint64_t)(p.lmax ? p.lmax - 1 : 0), p.n);
      for (uint64_t e = p.lead + lane; e < last && e < kEnds; e += 32)
        mark (e, ACM_S2_DMASK_ALL);
    }
    for (uint64_t s = span_base + 2 * (uint64_t)lane; s < span_end; s += 64)
      if (s >= 2)
        s2_probe (p, text8, s, [&] (uint32_t d) {
          const uint64_t e = s + d;
          if (e >= p.lead && e < p.n)
            mark (e, 1u << (d >> 1));
        });
    __syncwarp ();
    uint32_t total = 0, spill = 0;
    for (uint32_t w = 0; w < kEnds; w += 32) {
      const uint32_t rel = w + lane;
      const bool on = ((s_mask32[rel >> 1] >> (16 * (rel & 1u))) & 0xFFFFu) != 0;
      const uint32_t m = __ballot_sync (kFull, on);
      total += __popc (m);
      if (w >= kS2SpanBytes)
        spill += __popc (m);
    }
    unsigned long long seg = 0;
    if (lane == 0 && total)
      seg = atomicAdd (p.cand_count, (unsigned long long)total);
    seg = __shfl_sync (kFull, seg, 0);
    if (seg + total > p.cand_cap) {
      if (lane == 0)
        atomicExch (p.overflow, 1u);
      total = spill = 0;
    }
    uint64_t at = seg;
    if (total)
      for (uint32_t w = 0; w < kEnds; w += 32) {
        const uint32_t rel = w + lane;
        const uint32_t dmask = (s_mask32[rel >> 1] >> (16 * (rel & 1u))) & 0xFFFFu;
        const uint32_t m = __ballot_sync (kFull, dmask != 0);
        if (dmask)
          p.cand_pos[at + __popc (m & ((1u << lane) - 1u))] = (span_base + rel) | ((uint64_t)dmask << 48);
        at += __popc (m);
      }
    if (lane == 0) {
      p.tile_first[span] = seg;
      p.tile_n[span] = total;
      p.tile_spill[span] = spill;
    }
    __syncwarp ();
  }
}

/* F2 / F4: one thread per candidate walks the reverse trie leftwards from the candidate's position.
 * F2 (kEmit = false) counts the keywords ending there; F3 sums the counts of each tile (its candidates are contiguous and in
 * position order); after the device scan of the tile totals, F4 (kEmit = true) walks again and writes the records of the
 * candidate at tile offset + the counts of the candidates before it in the tile, longest keyword first.
 * The grid is fixed and the number of candidates is read from device memory: no host round trip before these kernels. */
template <int W, bool kEmit>
__global__ void __launch_bounds__ (256)
filter_verify_kernel (const __grid_constant__ FilterParams p) {
  const uint64_t nb_candidates = min ((uint64_t)*p.cand_count, p.cand_cap);
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nb_candidates; c += (uint64_t)gridDim.x * blockDim.x) {
  uint32_t expected = 0;
  uint64_t out = 0;
  if (kEmit) { /* a candidate without matches costs F4 this one load */
    expected = p.cand_matches[c];
    if (!expected)
      continue;
  }
  const uint64_t packed = p.cand_pos[c];
  const int64_t pos = (int64_t)(packed & ACM_CAND_POS_MASK);
  const uint32_t dmask = (uint32_t)(packed >> 48); /* 0: every keyword ending here is reported */
  if (kEmit) {
    const uint64_t tile = (uint64_t)pos / p.tile_syms;
    out = p.tile_offsets[tile] + p.cand_prefix[c];
    if (expected <= 2) { /* usual case: the counting pass kept them, no second walk */
      const uint4 m = p.cand_inline[c];
      if (expected == 2) {
        if (out < p.capacity)
          p.matches[out] = ACMB200Match{ p.base + (uint64_t)pos, m.z, m.w };
        out++;
      }
      if (out < p.capacity)
        p.matches[out] = ACMB200Match{ p.base + (uint64_t)pos, m.x, m.y };
      continue;
    }
  }
  uint4 first_two = make_uint4 (0, 0, 0, 0);
  uint64_t key;
  uint32_t node, kw, found = 0;
  auto report = [&] (uint32_t keyword, uint32_t len) {
    if (dmask) { /* stride-2 candidates: only the keywords whose chosen window produced this candidate */
      const uint32_t dd = p.kw_dist[keyword], d = (pos & 1) ? dd >> 8 : dd & 0xFFu;
      if (!((dmask >> (d >> 1)) & 1u))
        return;
    }
    if (kEmit) { /* found shortest first; the record order is longest first */
      const uint64_t at = out + (expected - 1 - found);
      if (at < p.capacity)
        p.matches[at] = ACMB200Match{ p.base + (uint64_t)pos, keyword, len };
    } else if (found == 0) {
      first_two.x = keyword;
      first_two.y = len;
    } else if (found == 1) {
      first_two.z = keyword;
      first_two.w = len;
    }
    found++;
  };
  if (qgram_key_at<W> (p, pos, &key) && slot_lookup (p.qgrams, p.qgram_mask, key, &node, &kw)) {
    uint32_t len = p.q;
    for (;;) {
      if (node != ACM_TAB_NONE && (node & ACM_TAIL_FLAG)) {
        /* exactly one keyword lies below: compare its remaining symbols with the text, right to left */
        const uint32_t k = node & ~ACM_TAIL_FLAG;
        uint32_t klen;
        bool same = true;
        if (W == 1 && p.kw_meta && (len & 3u) == 0 && (uint64_t)pos + 1 >= __ldg (&p.kw_meta[k]).x) {
          /* the keyword lies inside the text: four bytes per step, aligned words of the text against the keyword's reversed,
           * word-aligned copy.  Step i compares text bytes e-3..e (e = pos - len - 4i) with reversed-pool word len/4 + i. */
          const uint2 meta = __ldg (&p.kw_meta[k]);
          klen = meta.x;
          const uint32_t *rp = p.kw_rpool + meta.y;
          const uint32_t *tw = reinterpret_cast<const uint32_t *> (p.text);
          if (len < klen) {
            int64_t e = pos - (int64_t)len, wi = e >> 2;
            const uint32_t sh = ((uint32_t)(e & 3) + 1u) * 8u;
            uint32_t hi = tw[wi];
            for (uint32_t j = len; j < klen && same; j += 8, wi -= 2) { /* two steps per round: their four loads are in flight together */
              const bool two = j + 4 < klen;
              const uint32_t lo1 = wi > 0 ? tw[wi - 1] : 0u, lo2 = two && wi > 1 ? tw[wi - 2] : 0u;
              const uint32_t r1 = __ldg (rp + (j >> 2)), r2 = two ? __ldg (rp + (j >> 2) + 1) : 0u;
              const uint32_t t1 = __byte_perm (__funnelshift_rc (lo1, hi, sh), 0u, 0x0123); /* lowest byte = text byte e */
              const uint32_t t2 = __byte_perm (__funnelshift_rc (lo2, lo1, sh), 0u, 0x0123);
              const uint32_t rem = klen - j, m1 = rem >= 4 ? 0xFFFFFFFFu : (1u << (8 * rem)) - 1u;
              const uint32_t m2 = !two ? 0u : (rem >= 8 ? 0xFFFFFFFFu : (1u << (8 * (rem - 4))) - 1u);
              same = (((t1 ^ r1) & m1) | ((t2 ^ r2) & m2)) == 0;
              hi = lo2;
            }
          }
        } else {
          klen = p.kw_len[k];
          const typename SymT<W>::type *kwsym = reinterpret_cast<const typename SymT<W>::type *> (p.kw_pool) + p.kw_off[k];
          if ((uint64_t)pos + 1 >= klen) {
            /* the keyword lies inside the text: four symbols per step, all eight loads of a step issued before the first
             * compare (one symbol per step was one L2 round trip per symbol) */
            const typename SymT<W>::type *tx = reinterpret_cast<const typename SymT<W>::type *> (p.text) + pos;
            for (uint32_t j = len; j < klen && same; j += 4) {
              uint32_t diff = 0;
#pragma unroll
              for (uint32_t u = 0; u < 4; u++)
                if (j + u < klen)
                  diff |= (uint32_t)*(tx - (int64_t)(j + u)) ^ (uint32_t)kwsym[klen - 1 - j - u];
              same = diff == 0;
            }
          } else
            for (uint32_t j = len; j < klen && same; j++) {
              uint32_t sym;
              same = symbol_at<W> (p, pos - (int64_t)j, &sym) && sym == kwsym[klen - 1 - j];
            }
        }
        if (same)
          report (k, klen);
        break;
      }
      if (kw != ACM_TAB_NONE)
        report (kw, len);
      uint32_t sym;
      if (!symbol_at<W> (p, pos - (int64_t)len, &sym))
        break;
      if (!slot_lookup (p.edges, p.edge_mask, ((uint64_t)node << 32) | sym, &node, &kw))
        break;
      len++;
    }
  }
  if (!kEmit) {
    p.cand_matches[c] = found;
    p.cand_inline[c] = first_two;
  }
  }
}

/* F3: matches per tile = sum over the candidates that END in it, in end order -- its own ones (minus those the stride-2 kernel
 * found ending in the next span) merged with the ones the previous span left for it -- and the running prefix of every such
 * candidate, i.e. where F4 writes its records relative to the tile's offset. */
__global__ void __launch_bounds__ (256)
filter_tile_totals_kernel (const __grid_constant__ FilterParams p) {
  const uint64_t tile = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= p.ntiles)
    return;
  const uint32_t n_main = p.tile_n[tile] - (p.tile_spill ? p.tile_spill[tile] : 0u);
  const uint64_t first = p.tile_first[tile];
  uint32_t n_prev = 0;
  uint64_t first_prev = 0;
  if (p.tile_spill && tile) {
    n_prev = p.tile_spill[tile - 1];
    first_prev = p.tile_first[tile - 1] + p.tile_n[tile - 1] - n_prev;
  }
  uint32_t total = 0, i = 0, j = 0;
  while (i < n_prev || j < n_main) {
    bool take_prev = i < n_prev;
    if (take_prev && j < n_main)
      take_prev = (p.cand_pos[first_prev + i] & ACM_CAND_POS_MASK) <= (p.cand_pos[first + j] & ACM_CAND_POS_MASK);
    const uint64_t c = take_prev ? first_prev + i++ : first + j++;
    p.cand_prefix[c] = total;
    total += p.cand_matches[c];
  }
  p.tile_matches[tile] = total;
}

/* F3 of the dense mode (no spill lists, tens of candidates per tile): one WARP per tile, 32 candidates per step, coalesced; the
 * thread-per-tile kernel above walked each tile's list serially (0.22 ms of a 1.5 ms config-5 step). */
__global__ void __launch_bounds__ (256)
filter_tile_totals_dense_kernel (const __grid_constant__ FilterParams p) {
  const int lane = threadIdx.x & 31;
  const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  for (uint64_t tile = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < p.ntiles; tile += nwarps) {
    const uint32_t n = p.tile_n[tile];
    const uint64_t first = p.tile_first[tile];
    uint32_t total = 0, v_next = (uint32_t)lane < n ? p.cand_matches[first + lane] : 0u;
    for (uint32_t i = 0; i < n; i += 32) {
      const bool have = i + lane < n;
      const uint32_t v = v_next;
      v_next = i + 32 + lane < n ? p.cand_matches[first + i + 32 + lane] : 0u; /* in flight during the scan below */
      uint32_t incl = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync (kFull, incl, d);
        if (lane >= d)
          incl += o;
      }
      if (have)
        p.cand_prefix[first + i + lane] = total + incl - v;
      total += __shfl_sync (kFull, incl, 31);
    }
    if (lane == 0)
      p.tile_matches[tile] = total;
  }
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* In-place table update after append-only insertions (acm_patch_filter_tables): one thread per changed word / slot.    */
/* ------------------------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__ (256)
apply_patches_kernel (const acm_patch *__restrict__ patches, uint64_t nb, uint32_t *bloom, uint32_t *bloom2, acm_slot *qgrams, uint32_t *qset, acm_slot *edges) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb)
    return;
  const acm_patch p = patches[i];
  switch (p.array) {
    case ACM_PATCH_BLOOM:
      bloom[p.index] = p.value[0];
      break;
    case ACM_PATCH_BLOOM2:
      bloom2[p.index] = p.value[0];
      break;
    case ACM_PATCH_QSET:
      qset[p.index] = p.value[0];
      break;
    case ACM_PATCH_QGRAMS:
      *reinterpret_cast<uint4 *> (qgrams + p.index) = make_uint4 (p.value[0], p.value[1], p.value[2], p.value[3]);
      break;
    case ACM_PATCH_EDGES:
      *reinterpret_cast<uint4 *> (edges + p.index) = make_uint4 (p.value[0], p.value[1], p.value[2], p.value[3]);
      break;
  }
}

} // namespace acm
