/* acm_device.cu -- GPU side of libac75.so: device images, scan pipelines, batch C-ABI (include/acm_b200.h).
 *
 * There is deliberately no CPU implementation of the batch scan in this library: without a CUDA device every entry point
 * returns ACM_B200_ERR_NO_DEVICE.
 */
#include "acm_internal.h"
#include "acm_kernels.cuh"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace acm;

static thread_local char g_error[512] = "";

static int
fail (int code, const char *fmt, const char *a = "", const char *b = "") {
  snprintf (g_error, sizeof g_error, fmt, a, b);
  return code;
}

#define CUDA_TRY(call)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return fail (ACM_B200_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString (e_));     \
  } while (0)

extern "C" const char *
acm_b200_last_error (void) {
  return g_error;
}

extern "C" const char *
acm_b200_version (void) {
  return "ac75-b200 0.1 (sm_100a)";
}

extern "C" int
acm_b200_device_count (void) {
  int n = 0;
  if (cudaGetDeviceCount (&n) != cudaSuccess) {
    cudaGetLastError ();
    return 0;
  }
  return n;
}

/* grow-only device buffer */
struct DevBuf {
  void *ptr = nullptr;
  size_t bytes = 0;
  int ensure (size_t want) {
    if (want <= bytes)
      return ACM_B200_OK;
    if (ptr)
      cudaFree (ptr);
    ptr = nullptr;
    bytes = 0;
    size_t grown = want + want / 8 + 256;
    cudaError_t e = cudaMalloc (&ptr, grown);
    if (e != cudaSuccess) {
      cudaGetLastError ();
      e = cudaMalloc (&ptr, grown = want);
    }
    if (e != cudaSuccess)
      return fail (ACM_B200_ERR_NOMEM, "cudaMalloc of %s bytes failed: %s", std::to_string (want).c_str (), cudaGetErrorString (e));
    bytes = grown;
    return ACM_B200_OK;
  }
  void release () {
    if (ptr)
      cudaFree (ptr);
    ptr = nullptr;
    bytes = 0;
  }
  template <typename T> T *as () const { return reinterpret_cast<T *> (ptr); }
};

/* Everything ONE scan in flight needs besides the tables: streams, events, scratch buffers that grow on demand and are reused, the
 * pinned copy of the scalars the kernels write.  A machine keeps a pool of these, so several host threads can scan one machine at
 * the same time (the reference's scan is lock-free with a caller-owned cursor, README.md:266,364): the machine lock is held only to
 * pick a context and to merge the statistics, never while the GPU works. */
struct ScanContext {
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev[6] = {}, ev_copy[4] = {};
  DevBuf d_text, d_text2, d_matches, d_matches2, d_counts, d_offsets, d_block_sums, d_cand_pos, d_cand_matches, d_cand_prefix, d_cand_inline, d_tile_first, d_tile_n, d_tile_spill, d_hot_spans, d_events, d_chunk_events, d_small;
  struct Small { /* one pinned + one device copy of the scalars the kernels write */
    unsigned long long cand_count;
    uint64_t grand_total;
    uint32_t overflow;
    uint32_t pad;
    unsigned long long span_counter;
    unsigned int hot_count; /* spans the stride-2 kernel left to filter_hot_spans_kernel */
    uint32_t pad2;
    uint32_t prefix[1024];
  } *h_small = nullptr;
  ACMB200Stats stats = {}; /* of the scan this context is running: merged into the machine's under the lock */
  bool in_use = false;

  int init () {
    CUDA_TRY (cudaStreamCreateWithFlags (&stream, cudaStreamNonBlocking));
    CUDA_TRY (cudaStreamCreateWithFlags (&copy_stream, cudaStreamNonBlocking));
    for (cudaEvent_t &e : ev)
      CUDA_TRY (cudaEventCreate (&e));
    for (cudaEvent_t &e : ev_copy)
      CUDA_TRY (cudaEventCreate (&e));
    CUDA_TRY (cudaMallocHost (&h_small, sizeof (*h_small)));
    return d_small.ensure (sizeof (*h_small));
  }
  void release () {
    for (DevBuf *b : { &d_text, &d_text2, &d_matches, &d_matches2, &d_counts, &d_offsets, &d_block_sums, &d_cand_pos, &d_cand_matches, &d_cand_prefix, &d_cand_inline, &d_tile_first,
                       &d_tile_n, &d_tile_spill, &d_hot_spans, &d_events, &d_chunk_events, &d_small })
      b->release ();
    for (cudaEvent_t e : ev)
      if (e)
        cudaEventDestroy (e);
    for (cudaEvent_t e : ev_copy)
      if (e)
        cudaEventDestroy (e);
    if (stream)
      cudaStreamDestroy (stream);
    if (copy_stream)
      cudaStreamDestroy (copy_stream);
    if (h_small)
      cudaFreeHost (h_small);
  }
};

/* The tables of one generation of the dictionary on one device.  Immutable while a scan uses it (refs > 0): a finalise that finds
 * the dictionary changed while scans are in flight builds a new image and retires this one. */
struct acm_device_image {
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  acm_tables tab = {}; /* host images; big arrays are freed after upload except dfa_of_state */
  DevBuf d_delta, d_out_offsets, d_out_entries, d_bloom, d_bloom2, d_bloom_s2, d_s2_dist, d_kw_dist, d_qgrams, d_qset, d_edges, d_kw_len, d_kw_off, d_kw_pool, d_kw_meta, d_kw_rpool, d_patches;
  bool two_level = false; /* the filter engine uses the second-level filter in global memory */
  bool stride2 = false;   /* the stride-2 tables (bloom_s2, s2_dist, kw_dist) are resident */
  bool has_rpool = false; /* the reversed keyword pool (kw_meta, kw_rpool) is resident */
  volatile bool prefer_dense = false; /* the last filter scan overflowed its candidate buffers: start the next one in dense mode */
  volatile double cand_rate = -1;     /* candidates per symbol seen by the last filter scan (sizes the candidate list); < 0: none yet */
  uint64_t generation = 0; /* of the dictionary these tables were built from */
  int refs = 0;            /* scans in flight on this image (machine lock) */
  std::vector<ScanContext *> contexts;
  struct acm_device_image *retired = nullptr; /* older images still used by a scan in flight */
  ACMB200Stats stats = {};
};

static void
free_image (acm_device_image *img) {
  for (DevBuf *b : { &img->d_delta, &img->d_out_offsets, &img->d_out_entries, &img->d_bloom, &img->d_bloom2, &img->d_bloom_s2, &img->d_s2_dist, &img->d_kw_dist, &img->d_qgrams, &img->d_qset,
                     &img->d_edges, &img->d_kw_len, &img->d_kw_off, &img->d_kw_pool, &img->d_kw_meta, &img->d_kw_rpool, &img->d_patches })
    b->release ();
  for (ScanContext *cx : img->contexts) {
    cx->release ();
    delete cx;
  }
  acm_free_tables (&img->tab);
  delete img;
}

extern "C" void
acm_device_release (struct acm_device_image *img) {
  if (!img)
    return;
  int prev = -1;
  cudaGetDevice (&prev);
  cudaSetDevice (img->device);
  while (img) {
    acm_device_image *older = img->retired;
    free_image (img);
    img = older;
  }
  if (prev >= 0)
    cudaSetDevice (prev);
}

/* restores the caller's current device when an entry point returns */
struct DeviceGuard {
  int prev = -1;
  DeviceGuard () { if (cudaGetDevice (&prev) != cudaSuccess) { cudaGetLastError (); prev = -1; } }
  ~DeviceGuard () { if (prev >= 0) cudaSetDevice (prev); }
};

static int
upload (DevBuf &dst, const void *src, size_t bytes, cudaStream_t st, size_t reserve = 0) {
  int rc = dst.ensure (std::max (reserve, bytes ? (bytes + 15) / 16 * 16 : (size_t)16));
  if (rc)
    return rc;
  if (bytes)
    CUDA_TRY (cudaMemcpyAsync (dst.ptr, src, bytes, cudaMemcpyHostToDevice, st));
  return ACM_B200_OK;
}

/* ---- finalise ------------------------------------------------------------------------------------------------------- */
/* uploads the host images of img->tab and frees the big ones */
static int
upload_tables (acm_device_image *img, cudaStream_t st, uint64_t *table_bytes) {
  acm_tables &t = img->tab;
  int rc;
  uint64_t bytes = 0;
  if (t.engine == ACM_B200_ENGINE_FILTER) {
    if ((rc = upload (img->d_bloom, t.bloom, (size_t)t.bloom_words * 4, st)) || (t.bloom2 && (rc = upload (img->d_bloom2, t.bloom2, (size_t)t.bloom2_words * 4, st))) || (rc = upload (img->d_qgrams, t.qgrams, t.qgram_slots * sizeof (acm_slot), st))
        || (rc = upload (img->d_edges, t.edges, t.edge_slots * sizeof (acm_slot), st))
        || (t.qset && (rc = upload (img->d_qset, t.qset, (size_t)16 << (32 - t.qset_shift), st)))
        || (t.bloom_s2 && ((rc = upload (img->d_bloom_s2, t.bloom_s2, (size_t)t.bloom_s2_words * 4, st)) || (rc = upload (img->d_s2_dist, t.s2_dist, (size_t)4 << t.s2_dist_log2, st))
                            || (rc = upload (img->d_kw_dist, t.kw_dist, ((size_t)t.nb_keywords + 1) * 2, st))))
        || (rc = upload (img->d_kw_len, t.kw_len, ((size_t)t.nb_keywords + 1) * 4, st, t.builder ? t.builder->kw_cap * 4 : 0))
        || (rc = upload (img->d_kw_off, t.kw_off, ((size_t)t.nb_keywords + 1) * 8, st, t.builder ? t.builder->kw_cap * 8 : 0)) || (rc = upload (img->d_kw_pool, t.kw_pool, t.kw_pool_bytes, st))
        || (t.kw_meta && ((rc = upload (img->d_kw_meta, t.kw_meta, ((size_t)t.nb_keywords + 1) * 8, st)) || (rc = upload (img->d_kw_rpool, t.kw_rpool, t.kw_rpool_words * 4, st)))))
      return rc;
    bytes = (uint64_t)t.bloom_words * 4 + (t.qgram_slots + t.edge_slots) * sizeof (acm_slot) + t.kw_pool_bytes + (uint64_t)t.nb_keywords * 12
            + (t.qset ? (uint64_t)16 << (32 - t.qset_shift) : 0) + (t.bloom_s2 ? (uint64_t)t.bloom_s2_words * 4 + ((uint64_t)4 << t.s2_dist_log2) + (uint64_t)t.nb_keywords * 2 : 0) + (t.kw_meta ? (uint64_t)t.nb_keywords * 8 + t.kw_rpool_words * 4 : 0);
  } else {
    if ((rc = upload (img->d_delta, t.delta, t.delta_bytes, st)) || (rc = upload (img->d_out_offsets, t.out_offsets, ((size_t)t.nb_dfa_states - t.out_threshold + 1) * 4, st))
        || (rc = upload (img->d_out_entries, t.out_entries, t.nb_out_entries * sizeof (acm_output), st)))
      return rc;
    bytes = t.delta_bytes + ((uint64_t)t.nb_dfa_states - t.out_threshold + 1) * 4 + t.nb_out_entries * sizeof (acm_output);
  }
  CUDA_TRY (cudaStreamSynchronize (st));
  img->two_level = t.bloom2 != nullptr;
  img->stride2 = t.bloom_s2 != nullptr;
  img->has_rpool = t.kw_meta != nullptr;
  img->prefer_dense = false;
  img->cand_rate = -1;
  *table_bytes = bytes;
  if (t.builder) /* kept for in-place updates after append-only insertions: the host images stay */
    return ACM_B200_OK;
  /* the big host images are not needed any more */
  free (t.delta), t.delta = nullptr;
  free (t.out_offsets), t.out_offsets = nullptr;
  free (t.out_entries), t.out_entries = nullptr;
  free (t.bloom), t.bloom = nullptr;
  free (t.bloom2), t.bloom2 = nullptr;
  free (t.bloom_s2), t.bloom_s2 = nullptr;
  free (t.s2_dist), t.s2_dist = nullptr;
  free (t.kw_dist), t.kw_dist = nullptr;
  free (t.qgrams), t.qgrams = nullptr;
  free (t.qset), t.qset = nullptr;
  free (t.kw_len), t.kw_len = nullptr;
  free (t.kw_off), t.kw_off = nullptr;
  free (t.kw_pool), t.kw_pool = nullptr;
  free (t.kw_meta), t.kw_meta = nullptr;
  free (t.kw_rpool), t.kw_rpool = nullptr;
  free (t.edges), t.edges = nullptr;
  return ACM_B200_OK;
}

/* Uploads what an in-place update changed: the appended tails of the per-keyword arrays as contiguous copies, the scattered words
 * and slots as a patch list applied by one small kernel. */
static int
apply_patches (acm_device_image *img, const acm_patch_list &pl) {
  const acm_tables &t = img->tab;
  cudaStream_t st = cudaStreamPerThread;
  const size_t w = (size_t)t.width;
  if ((pl.kw_first + pl.kw_nb + 1) * 8 > img->d_kw_off.bytes || (pl.kw_first + pl.kw_nb + 1) * 4 > img->d_kw_len.bytes || (pl.pool_first_sym + pl.pool_nb_syms) * w > img->d_kw_pool.bytes)
    return fail (ACM_B200_ERR_NOMEM, "in-place update: device arrays too small%s", "");
  if (pl.kw_nb) {
    CUDA_TRY (cudaMemcpyAsync (img->d_kw_len.as<uint32_t> () + pl.kw_first, t.kw_len + pl.kw_first, pl.kw_nb * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY (cudaMemcpyAsync (img->d_kw_off.as<uint64_t> () + pl.kw_first, t.kw_off + pl.kw_first, pl.kw_nb * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY (cudaMemcpyAsync (img->d_kw_pool.as<unsigned char> () + pl.pool_first_sym * w, reinterpret_cast<const unsigned char *> (t.kw_pool) + pl.pool_first_sym * w, pl.pool_nb_syms * w,
                               cudaMemcpyHostToDevice, st));
  }
  if (pl.nb) {
    int rc = img->d_patches.ensure (pl.nb * sizeof (acm_patch));
    if (rc)
      return rc;
    CUDA_TRY (cudaMemcpyAsync (img->d_patches.ptr, pl.items, pl.nb * sizeof (acm_patch), cudaMemcpyHostToDevice, st));
    apply_patches_kernel<<<(unsigned)((pl.nb + 255) / 256), 256, 0, st>>> (img->d_patches.as<acm_patch> (), pl.nb, img->d_bloom.as<uint32_t> (), img->d_bloom2.as<uint32_t> (), img->d_qgrams.as<acm_slot> (),
                                                                          img->d_qset.as<uint32_t> (), img->d_edges.as<acm_slot> ());
    CUDA_TRY (cudaGetLastError ());
  }
  CUDA_TRY (cudaStreamSynchronize (st));
  return ACM_B200_OK;
}

/* Machine lock held.  Leaves in m->device an image of the CURRENT dictionary on `device` (-1: the current device, or the one the
 * machine is already on).  An image that a scan in flight is using is never touched: it is retired and a new one built. */
static int
finalise_locked (ACMachine *m, int device) {
  if (acm_b200_device_count () <= 0)
    return fail (ACM_B200_ERR_NO_DEVICE, "no CUDA device is available: the batch scan has no CPU fallback");
  acm_device_image *img = m->device;
  if (img && (device < 0 || img->device == device) && img->generation == m->generation) {
    CUDA_TRY (cudaSetDevice (img->device));
    return ACM_B200_OK;
  }
  if (device < 0) {
    if (img)
      device = img->device;
    else
      CUDA_TRY (cudaGetDevice (&device));
  }
  CUDA_TRY (cudaSetDevice (device));
  const auto t0 = std::chrono::steady_clock::now ();
  /* reuse the image (its device buffers) when nobody is scanning with it and it lives on the right device */
  const bool reuse = img && img->device == device && img->refs == 0;
  /* append-only insertions since these tables were built (Meyer-style updates between two scans): patch them in place */
  if (reuse && img->tab.builder && !m->force_rebuild && !m->option_no_patch) {
    acm_patch_list pl;
    if (acm_patch_filter_tables (m, &img->tab, &pl) == ACM_B200_OK) {
      int prc = apply_patches (img, pl);
      free (pl.items);
      if (prc)
        return prc;
      img->generation = m->generation;
      m->device_generation = m->generation;
      ACMB200Stats &s = img->stats;
      s.nb_states = img->tab.nb_states;
      s.nb_keywords = img->tab.nb_keywords;
      s.max_keyword_length = img->tab.lmax;
      s.finalise_count++;
      s.patch_count++;
      s.finalise_ms = std::chrono::duration<double, std::milli> (std::chrono::steady_clock::now () - t0).count ();
      return ACM_B200_OK;
    }
    g_error[0] = 0;
  }
  acm_device_image *fresh = reuse ? img : new acm_device_image ();
  auto abandon = [&] (int rc) {
    if (!reuse)
      free_image (fresh);
    else { /* a half-rebuilt image must not be scanned with: drop it */
      m->device = fresh->retired;
      fresh->retired = nullptr;
      free_image (fresh);
    }
    return rc;
  };
  if (!reuse) {
    fresh->device = device;
    int v = 0;
    if (cudaDeviceGetAttribute (&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess)
      return abandon (fail (ACM_B200_ERR_CUDA, "cudaDeviceGetAttribute failed%s", ""));
    fresh->sm_count = v;
    if (cudaDeviceGetAttribute (&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess)
      return abandon (fail (ACM_B200_ERR_CUDA, "cudaDeviceGetAttribute failed%s", ""));
    fresh->smem_optin = (size_t)v;
    if (img)
      fresh->stats = img->stats; /* counters are per machine */
  }
  acm_free_tables (&fresh->tab);
  /* shared-memory budget for the resident table: everything a block may opt in to, minus class map / staging / slack */
  const uint64_t budget = fresh->smem_optin - 32 * (192 * 2 + 16) - 2048;
  /* the stride-2 kernel works best when it leaves part of the SM's 256 KB to the L1 cache (its confirmation step re-reads text the
   * tile loads brought in): by default it takes the 196 KB carve-out, not the largest one */
  const uint64_t s2_smem = m->option_no_stride2 ? 0 : (m->option_s2_smem_kb ? std::min<uint64_t> (fresh->smem_optin, m->option_s2_smem_kb * 1024 - 1024) : std::min<uint64_t> (fresh->smem_optin, 195 * 1024));
  m->last_smem_budget = budget;
  m->last_s2_smem = s2_smem;
  int rc;
  bool from_blob = false;
  if (m->preloaded && m->preloaded_generation == m->generation && m->preloaded_budget <= budget && m->preloaded_s2_smem <= fresh->smem_optin && !m->engine_override[0]) {
    fresh->tab = *m->preloaded; /* the images of a blob (acm_blob.c): nothing to build */
    free (m->preloaded);
    m->preloaded = nullptr;
    from_blob = true;
  } else {
    acm_ensure_trie_locked (m);
    if ((rc = acm_build_tables (m, &fresh->tab, budget, s2_smem)))
      return abandon (fail (rc, "building the automaton tables failed%s", ""));
  }
  /* the upload runs on a stream of its own context-free: the legacy default stream of this thread */
  uint64_t bytes = 0;
  if ((rc = upload_tables (fresh, cudaStreamPerThread, &bytes)))
    return abandon (rc);
  const acm_tables &t = fresh->tab;
  fresh->generation = m->generation;
  if (!reuse) {
    fresh->retired = img; /* still referenced by scans in flight (or on another device): freed when the last one ends */
    if (img && img->refs == 0) {
      fresh->retired = img->retired;
      img->retired = nullptr;
      int prev = device;
      cudaSetDevice (img->device);
      free_image (img);
      cudaSetDevice (prev);
    }
    m->device = fresh;
  }
  m->device_generation = m->generation;
  m->force_rebuild = 0;
  ACMB200Stats &s = fresh->stats;
  s.engine = t.engine;
  s.symbol_width = t.width;
  s.nb_states = t.nb_states;
  s.nb_keywords = t.nb_keywords;
  s.max_keyword_length = t.lmax;
  s.min_keyword_length = t.lmin;
  s.nb_classes = t.nb_classes;
  s.table_bytes = bytes;
  s.filter_fp = t.engine == ACM_B200_ENGINE_FILTER ? (fresh->stride2 ? t.bloom_s2_hit_rate : t.bloom_fp) : 0;
  s.filter_stride = 0;
  s.finalise_count++;
  s.blob_loads += from_blob;
  s.finalise_ms = std::chrono::duration<double, std::milli> (std::chrono::steady_clock::now () - t0).count ();
  return ACM_B200_OK;
}

extern "C" int
acm_b200_finalise (ACMachine *m, int device) {
  if (!m)
    return fail (ACM_B200_ERR_INVALID, "null machine%s", "");
  DeviceGuard guard;
  acm_lock (m);
  int rc = finalise_locked (m, device);
  acm_unlock (m);
  return rc;
}

extern "C" int
acm_b200_set_option (ACMachine *m, const char *key, const char *value) {
  if (!m || !key || !value)
    return ACM_B200_ERR_INVALID;
  acm_lock (m);
  int rc = ACM_B200_OK;
  if (!strcmp (key, "engine")) {
    snprintf (m->engine_override, sizeof m->engine_override, "%s", strcmp (value, "auto") ? value : "");
    m->generation++, m->force_rebuild = 1; /* forces a rebuild */
  } else if (!strcmp (key, "bloom_words"))
    m->option_bloom_words = strtoull (value, 0, 10), m->generation++, m->force_rebuild = 1;
  else if (!strcmp (key, "threads"))
    m->option_threads = strtoull (value, 0, 10);
  else if (!strcmp (key, "stream_bytes"))
    m->option_stream_bytes = strtoull (value, 0, 10);
  else if (!strcmp (key, "s2_smem_kb"))
    m->option_s2_smem_kb = strtoull (value, 0, 10), m->generation++, m->force_rebuild = 1;
  else if (!strcmp (key, "s2_dist_log2")) /* log2 of the largest distance table (in words) the stride-2 filter may use: default 23 = 32 MB */
    m->option_s2_dist_log2 = strtoull (value, 0, 10), m->generation++, m->force_rebuild = 1;
  else if (!strcmp (key, "dfa_events")) /* 0: pass 2 of the DFA engines always walks the text again */
    m->option_no_events = !strtoull (value, 0, 10);
  else if (!strcmp (key, "s2_batches")) /* batches of 32 hits the stride-2 kernel confirms at a time: 1, 2 (default) or 3 */
    m->option_s2_batches = strtoull (value, 0, 10);
  else if (!strcmp (key, "dfa_tma")) /* 0: the DFA count pass loads its chunks per thread instead of staging them through shared memory with TMA */
    m->option_no_tma = !strtoull (value, 0, 10);
  else if (!strcmp (key, "dfa_lean")) /* 1: the TMA-staged count pass records events only and the records are counted from the event lists afterwards
                                        * (default 0: it counts them itself -- the two forms measured within 3 % of each other, either way round, on different boxes) */
    m->option_lean = strtoull (value, 0, 10) != 0;
  else if (!strcmp (key, "patch")) /* 0: every insertion between two scans rebuilds the tables (the in-place update is the default) */
    m->option_no_patch = !strtoull (value, 0, 10), m->generation++, m->force_rebuild = 1;
  else if (!strcmp (key, "stride2")) /* 0: keep the one-test-per-position filter kernel even where the stride-2 one applies */
    m->option_no_stride2 = !strtoull (value, 0, 10), m->generation++, m->force_rebuild = 1;
  else
    rc = ACM_B200_ERR_INVALID;
  acm_unlock (m);
  return rc;
}

extern "C" int
acm_b200_get_stats (ACMachine *m, ACMB200Stats *stats) {
  if (!m || !stats)
    return ACM_B200_ERR_INVALID;
  acm_lock (m);
  if (m->device)
    *stats = m->device->stats;
  else
    memset (stats, 0, sizeof (*stats));
  acm_unlock (m);
  return ACM_B200_OK;
}

/* ---- TMA descriptor of the DFA count pass --------------------------------------------------------------------------------- */
/* The text as a 2-D byte tensor: `rows` rows of `chunk` bytes (row = chunk of the DFA walk), fetched in boxes of 32 rows x 32 bytes.
 * cuTensorMapEncodeTiled is a driver entry point: it is looked up at run time, so libac75.so does not link libcuda. */
static int
encode_chunk_tensor (CUtensorMap *map, const void *text, uint64_t chunk, uint64_t rows) {
  typedef CUresult (*encode_fn) (CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint ("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
      cudaGetLastError ();
      return ACM_B200_ERR_CUDA;
    }
    encode = reinterpret_cast<encode_fn> (fn);
  }
  const cuuint64_t dims[2] = { chunk, rows }, strides[1] = { chunk };
  const cuuint32_t box[2] = { kTmaStageBytes, 32 }, elem[2] = { 1, 1 };
  return encode (map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *> (text), dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? ACM_B200_OK
             : ACM_B200_ERR_CUDA;
}

/* ---- device scan of counts ------------------------------------------------------------------------------------------ */
static int
device_exclusive_scan (ScanContext *cx, const uint32_t *counts, uint64_t n, uint64_t *offsets, uint64_t *d_grand, cudaStream_t st) {
  if (n <= kScanSmall) {
    scan_small_kernel<<<1, kScanThreads, 0, st>>> (counts, n, offsets, d_grand);
    cx->stats.total_kernel_launches += 1;
    CUDA_TRY (cudaGetLastError ());
    return ACM_B200_OK;
  }
  const uint64_t nblocks = (n + kScanBlock - 1) / kScanBlock;
  int rc = cx->d_block_sums.ensure ((nblocks + 1) * 8);
  if (rc)
    return rc;
  uint64_t *sums = cx->d_block_sums.as<uint64_t> ();
  scan_block_sums_kernel<<<(unsigned)nblocks, kScanThreads, 0, st>>> (counts, n, sums);
  scan_spine_kernel<<<1, kScanThreads, 0, st>>> (sums, nblocks, d_grand);
  scan_apply_kernel<<<(unsigned)nblocks, kScanThreads, 0, st>>> (counts, n, sums, offsets);
  cx->stats.total_kernel_launches += 3;
  CUDA_TRY (cudaGetLastError ());
  return ACM_B200_OK;
}

struct ScanJob {
  const void *d_text;
  uint64_t n, lead, base;
  ACMB200Match *d_matches; /* device buffer of `capacity` records, or null while only counting */
  uint64_t capacity;
  uint32_t init_dfa_state;
  uint32_t prefix_len; /* symbols staged in h_small->prefix */
  cudaStream_t st;
};

/* ---- DFA pipeline --------------------------------------------------------------------------------------------------- */
template <typename Entry, bool kShared>
static int
run_dfa (ACMachine *m, acm_device_image *img, ScanContext *cx, ScanJob &job, uint64_t *total, bool matches_on_device, ACMB200Match *user_matches) {
  const acm_tables &t = img->tab;
  DfaParams p = {};
  p.text = reinterpret_cast<const uint8_t *> (job.d_text);
  p.n = job.n;
  p.lead = job.lead;
  p.base = job.base;
  p.warm = m->max_depth ? m->max_depth - 1 : 0;
  p.init_state = job.init_dfa_state;
  p.delta = img->d_delta.ptr;
  p.K = t.nb_classes;
  p.nb_states = t.nb_dfa_states;
  p.out_threshold = t.out_threshold;
  p.out_offsets = img->d_out_offsets.as<uint32_t> ();
  p.out_entries = img->d_out_entries.as<acm_output> ();
  p.nb_out_states = t.nb_dfa_states - t.out_threshold;
  memcpy (p.class_of_byte, t.class_of_byte, 256);

  const int threads = kShared ? 1024 : 256;
  size_t smem = 256 + (kShared ? ((size_t)t.nb_dfa_states * t.nb_classes * sizeof (Entry) + 15) / 16 * 16 : 0);
  /* pass 1: a uint16 records-per-state table after the delta table when it fits and no state has more than 65535 records */
  const size_t counts_bytes = ((size_t)p.nb_out_states * 2 + 15) / 16 * 16;
  p.counts_in_smem = t.max_out_records <= 0xFFFF && smem + counts_bytes + 1024 <= img->smem_optin && counts_bytes <= 32768;
  const size_t count_smem = smem + (p.counts_in_smem ? counts_bytes : 0);
  const int blocks_per_sm = kShared ? 1 : 8;
  uint64_t want_threads = m->option_threads ? m->option_threads : (uint64_t)img->sm_count * blocks_per_sm * threads;
  const uint64_t min_chunk = std::max<uint64_t> (256, (uint64_t)(4 * p.warm + 15) / 16 * 16);
  p.chunk = std::max<uint64_t> (min_chunk, ((job.n + want_threads - 1) / want_threads + 15) / 16 * 16);
  p.chunk = std::min<uint64_t> (p.chunk, std::max<uint64_t> (min_chunk, 1ull << 30)); /* the kernels keep chunk-relative positions in 32 bits */
  if (kShared && sizeof (Entry) == 2 && !m->option_threads)
    p.chunk = std::min<uint64_t> (p.chunk, std::max<uint64_t> (min_chunk, 32768)); /* positions inside a chunk fit 16 bits (pass 1 events); huge texts: several chunks per thread */
  /* pass 1 with the text staged through shared memory by the TMA unit (dfa_scan_tma_kernel): shared-memory engine whose tables leave
   * room for the stage buffers; chunks of whole 32-byte stages; the chunks that lie wholly inside the text are the rows of the tensor,
   * a last partial chunk goes through the walking kernel */
  const size_t tma_smem = kTmaCtaBytes + count_smem + 16; /* (+ the zero entry that ends the records-per-state table) */
  bool use_tma = kShared && sizeof (Entry) == 2 && p.counts_in_smem && tma_smem <= img->smem_optin && job.n < (1ull << 32) && !m->option_no_tma && !m->option_threads && p.warm < 4096;
  if (use_tma) {
    /* small chunks: the partial last chunk, which one thread of the walking kernel does after this pass, stays short */
    p.chunk = std::min<uint64_t> (p.chunk, std::max<uint64_t> (min_chunk, 2048));
    p.chunk = (p.chunk + 31) / 32 * 32;
    use_tma = job.n / p.chunk >= 1 && p.chunk >= (p.warm + 31) / 32 * 32 && ((uintptr_t)job.d_text & 15) == 0;
  }
  p.nchunks = (job.n + p.chunk - 1) / p.chunk;
  p.tma_chunks = use_tma ? job.n / p.chunk : 0;
  const unsigned grid = (unsigned)std::min<uint64_t> ((p.nchunks + threads - 1) / threads, (uint64_t)img->sm_count * blocks_per_sm);

  int rc;
  if ((rc = cx->d_counts.ensure (p.nchunks * 4)) || (rc = cx->d_offsets.ensure (p.nchunks * 8)))
    return rc;
  p.chunk_counts = cx->d_counts.as<uint32_t> ();
  p.chunk_offsets = cx->d_offsets.as<uint64_t> ();
  auto *d_small = cx->d_small.as<ScanContext::Small> ();

  /* pass 1 records its events when it can (shared-memory engine: 16-bit output-state index; chunk offsets of 16 bits; one event
   * slot per 4 symbols, i.e. as many bytes as the text): pass 2 then expands them instead of walking the text again */
  bool use_events = kShared && sizeof (Entry) == 2 && p.chunk <= 65536 && p.nb_out_states <= 65536 && !m->option_no_events;
  if (use_events) {
    p.events_per_chunk = (uint32_t)(p.chunk / 4);
    if (cx->d_events.ensure (p.nchunks * p.events_per_chunk * 4) || cx->d_chunk_events.ensure (p.nchunks * 4)) {
      use_events = false; /* no room: walk twice */
      g_error[0] = 0;
    }
  }
  if (use_events) {
    p.events = cx->d_events.as<uint32_t> ();
    p.chunk_events = cx->d_chunk_events.as<uint32_t> ();
    p.events_overflow = &cx->d_small.as<ScanContext::Small> ()->overflow;
    cx->h_small->overflow = 0;
    CUDA_TRY (cudaMemcpyAsync (&cx->d_small.as<ScanContext::Small> ()->overflow, &cx->h_small->overflow, 4, cudaMemcpyHostToDevice, job.st));
  }
  void (*count_k) (const DfaParams) = use_events ? dfa_scan_kernel<Entry, kShared, false, true> : dfa_scan_kernel<Entry, kShared, false, false>;
  /* pass 2: warp-cooperative emit when positions relative to a warp's 32 chunks fit 32 bits (always, short of absurd chunk sizes) */
  const bool coop = p.chunk * 32 < (1ull << 32);
  void (*emit_k) (const DfaParams) = coop ? dfa_emit_kernel<Entry, kShared> : dfa_scan_kernel<Entry, kShared, true, false>;
  const size_t emit_smem = smem + (coop ? sizeof (EmitWarpState) * (threads / 32) : 0);
  CUDA_TRY (cudaFuncSetAttribute (count_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)count_smem));
  CUDA_TRY (cudaFuncSetAttribute (emit_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emit_smem));
  cx->stats.smem_bytes = emit_smem;

  CUtensorMap tmap;
  if (use_tma && encode_chunk_tensor (&tmap, job.d_text, p.chunk, p.tma_chunks) != ACM_B200_OK) {
    use_tma = false; /* no TMA descriptor: the walking kernel does everything */
    p.tma_chunks = 0;
    g_error[0] = 0;
  }
  CUDA_TRY (cudaEventRecord (cx->ev[0], job.st));
  /* lean pass 1: events only, records counted from the event lists afterwards.  Needs 2 x class in a byte, and event lists that do
   * not cross a 4 GiB boundary (the kernel advances the low word of its event pointer only): lists of 2^k bytes aligned to their
   * size, or a buffer that has no such boundary inside */
  bool lean = false;
  if (use_tma && use_events && p.K <= 128 && m->option_lean) {
    const uint64_t list_bytes = (uint64_t)p.events_per_chunk * 4, ev0 = (uint64_t)(uintptr_t)p.events, ev1 = ev0 + p.nchunks * list_bytes - 1;
    lean = ((list_bytes & (list_bytes - 1)) == 0 && ev0 % list_bytes == 0) || (ev0 >> 32) == (ev1 >> 32);
  }
  if (use_tma) {
    void (*tma_k) (const DfaParams, const CUtensorMap) = lean ? dfa_scan_tma_lean_kernel : (use_events ? dfa_scan_tma_kernel<true> : dfa_scan_tma_kernel<false>);
    CUDA_TRY (cudaFuncSetAttribute (tma_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem));
    const unsigned tgrid = (unsigned)std::min<uint64_t> ((p.tma_chunks + 1023) / 1024, (uint64_t)img->sm_count);
    tma_k<<<tgrid, 1024, tma_smem, job.st>>> (p, tmap);
    CUDA_TRY (cudaGetLastError ());
    cx->stats.total_kernel_launches += 1;
    cx->stats.dfa_tma_scans++;
    if (lean) { /* records per chunk from the events just recorded */
      DfaParams pc = p;
      pc.nchunks = p.tma_chunks;
      dfa_count_events_kernel<<<(unsigned)std::min<uint64_t> ((pc.nchunks + 7) / 8, (uint64_t)img->sm_count * 16), 256, 0, job.st>>> (pc);
      CUDA_TRY (cudaGetLastError ());
      cx->stats.total_kernel_launches += 1;
      cx->stats.dfa_lean_scans++;
    }
  }
  if (p.tma_chunks < p.nchunks) { /* everything, or the partial last chunk */
    p.first_chunk = p.tma_chunks;
    const unsigned rgrid = (unsigned)std::min<uint64_t> ((p.nchunks - p.tma_chunks + threads - 1) / threads, (uint64_t)grid);
    count_k<<<rgrid, threads, count_smem, job.st>>> (p);
    CUDA_TRY (cudaGetLastError ());
    p.first_chunk = 0;
  }
  CUDA_TRY (cudaEventRecord (cx->ev[1], job.st));
  if ((rc = device_exclusive_scan (cx, p.chunk_counts, p.nchunks, cx->d_offsets.as<uint64_t> (), &d_small->grand_total, job.st)))
    return rc;
  CUDA_TRY (cudaMemcpyAsync (&cx->h_small->grand_total, &d_small->grand_total, 8, cudaMemcpyDeviceToHost, job.st));
  if (use_events)
    CUDA_TRY (cudaMemcpyAsync (&cx->h_small->overflow, &d_small->overflow, 4, cudaMemcpyDeviceToHost, job.st));
  CUDA_TRY (cudaStreamSynchronize (job.st));
  *total = cx->h_small->grand_total;
  cx->stats.main_kernel_launches += 1;
  cx->stats.total_kernel_launches += 1;
  if (use_events && cx->h_small->overflow) {
    use_events = false; /* a chunk met more output states than it has event slots: pass 2 walks */
    if (lean) { /* ... and the counts taken from the truncated lists are wrong: counted again by walking */
      void (*recount_k) (const DfaParams) = dfa_scan_kernel<Entry, kShared, false, false>;
      CUDA_TRY (cudaFuncSetAttribute (recount_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)count_smem));
      recount_k<<<grid, threads, count_smem, job.st>>> (p);
      CUDA_TRY (cudaGetLastError ());
      if ((rc = device_exclusive_scan (cx, p.chunk_counts, p.nchunks, cx->d_offsets.as<uint64_t> (), &d_small->grand_total, job.st)))
        return rc;
      CUDA_TRY (cudaMemcpyAsync (&cx->h_small->grand_total, &d_small->grand_total, 8, cudaMemcpyDeviceToHost, job.st));
      CUDA_TRY (cudaStreamSynchronize (job.st));
      *total = cx->h_small->grand_total;
      cx->stats.total_kernel_launches += 1;
    }
  }

  const uint64_t want = std::min<uint64_t> (*total, job.capacity);
  CUDA_TRY (cudaEventRecord (cx->ev[2], job.st));
  if (want) {
    if (matches_on_device)
      p.matches = user_matches;
    else {
      if ((rc = cx->d_matches.ensure (want * sizeof (ACMB200Match))))
        return rc;
      p.matches = cx->d_matches.as<ACMB200Match> ();
    }
    p.capacity = want;
    if (use_events)
      cx->stats.dfa_event_scans++;
    if (use_events)
      dfa_emit_events_kernel<<<(unsigned)std::min<uint64_t> ((p.nchunks + 7) / 8, (uint64_t)img->sm_count * 16), 256, 0, job.st>>> (p);
    else
      emit_k<<<grid, threads, emit_smem, job.st>>> (p);
    CUDA_TRY (cudaGetLastError ());
    cx->stats.main_kernel_launches += 1;
    cx->stats.total_kernel_launches += 1;
  }
  CUDA_TRY (cudaEventRecord (cx->ev[3], job.st));
  job.d_matches = p.matches;
  cx->stats.last_nb_candidates = 0;
  return ACM_B200_OK;
}

/* ---- filter pipeline ------------------------------------------------------------------------------------------------ */
enum { kFilterOverflow = 1000 }; /* internal: a candidate buffer overflowed, retry in dense mode */

/* One run of the filter pipeline over job.d_text[0 .. job.n): F1 (F1s + F1h), F2, F3, scan, F4.  Records go to out[0 .. out_cap).
 * Nothing between the first and the last kernel waits for the host: the verification kernels run on fixed grids and read the
 * number of candidates (and of unfinished spans) from device memory; the scalars come back once, after the last kernel.  Only a
 * scan into the library's own record buffer, which is sized from the total, synchronises before F4. */
template <int W>
static int
run_filter_once (ACMachine *m, acm_device_image *img, ScanContext *cx, const ScanJob &job, bool dense, uint64_t sparse_cand_cap, ACMB200Match *out, uint64_t out_cap, bool size_out_lazily,
                 uint64_t *total, ACMB200Match **out_used, bool first_segment, bool last_segment) {
  const acm_tables &t = img->tab;
  constexpr int kRowsOpt = 4; /* rows of 512 bytes per warp tile; 2 and 8 were measured slower */
  auto *d_small = cx->d_small.as<ScanContext::Small> ();
  FilterParams p = {};
  p.text = job.d_text;
  p.n = job.n;
  p.lead = job.lead;
  p.base = job.base;
  p.q = t.q;
  p.lmax = t.lmax;
  const bool s2 = W == 1 && !dense && img->stride2; /* stride-2 kernel: its "tiles" for F2..F4 are spans of 2 KiB tiles */
  /* dense mode stages every position of a tile (a stage as large as the tile): with 4-row tiles only 8 warps fit next to the filter
   * (12 % occupancy, issue 11 %: profiles/README.md round 2), so the dense mode takes small tiles with byte-sized stage entries where
   * they suffice: all 32 warps for 32- and 16-bit symbols, 8 for bytes */
  constexpr int kDenseRows = W == 4 ? 2 : 1; /* 256 / 256 / 512 symbols per dense tile for 32- / 16- / 8-bit symbols; up to 256 the stage entries are bytes */
  p.tile_syms = s2 ? kS2SpanBytes : (dense ? kDenseRows : kRowsOpt) * 512 / W;
  p.ntiles = (job.n + p.tile_syms - 1) / p.tile_syms;
  p.bloom_s2 = img->d_bloom_s2.as<uint32_t> ();
  p.bloom_s2_words = t.bloom_s2_words;
  p.s2_dist = img->d_s2_dist.as<uint32_t> ();
  p.s2_dist_log2 = t.s2_dist_log2;
  p.kw_dist = s2 ? img->d_kw_dist.as<uint16_t> () : nullptr;
  p.span_counter = &d_small->span_counter;
  p.s2_hit_cap = t.s2_hit_cap;
  p.hot_count = &d_small->hot_count;
  p.hot_cap = s2 ? (uint32_t)std::max<uint64_t> (p.ntiles / 8, 64) : 0; /* more unfinished spans than that: the text is dense, use the dense mode */
  p.bloom = img->d_bloom.as<uint32_t> ();
  p.bloom_words = t.bloom_words;
  p.bloom2 = img->two_level ? img->d_bloom2.as<uint32_t> () : nullptr;
  p.bloom2_words = t.bloom2_words;
  p.qgrams = img->d_qgrams.as<acm_slot> ();
  p.qgram_mask = t.qgram_slots - 1;
  p.qset = W == 4 ? nullptr : img->d_qset.as<uint4> ();
  p.qset_shift = t.qset_shift;
  p.qset_has_empty_key = t.qset_has_empty_key;
  p.edges = img->d_edges.as<acm_slot> ();
  p.edge_mask = t.edge_slots - 1;
  p.kw_len = img->d_kw_len.as<uint32_t> ();
  p.kw_off = img->d_kw_off.as<uint64_t> ();
  p.kw_pool = img->d_kw_pool.ptr;
  p.kw_meta = img->has_rpool ? img->d_kw_meta.as<uint2> () : nullptr;
  p.kw_rpool = img->d_kw_rpool.as<uint32_t> ();
  p.prefix = d_small->prefix;
  p.prefix_len = job.prefix_len;
  /* stage sized for the filter's expected raw hits per tile (false positives + a margin); the dense retry takes the whole tile */
  const uint32_t expected_hits = (uint32_t)(t.bloom_fp * p.tile_syms);
  p.stage_cap = dense ? p.tile_syms : std::min<uint32_t> (p.tile_syms, std::max<uint32_t> (192, 2 * expected_hits + 128));
  /* candidate list (32 bytes of scratch per entry): every position in dense mode; otherwise what the caller expects from the
   * candidate rate it has seen (run_filter) -- a denser text overflows it and is redone with a larger list or in dense mode */
  p.cand_cap = dense ? job.n : std::min<uint64_t> (job.n, sparse_cand_cap);
  size_t stage_bytes_per_warp = (size_t)p.stage_cap * (dense && p.tile_syms <= 256 ? 1 : 2) + 16; /* 16- or 8-bit positions (StageT of the kernel) + the warp's counter */
  int warps = 32;
  while (warps > 1 && (size_t)t.bloom_words * 4 + warps * stage_bytes_per_warp > img->smem_optin - 1024)
    warps /= 2;
  if (s2)
    warps = 32;
  const size_t smem = s2 ? (size_t)t.bloom_s2_words * 4 + 32 * (size_t)ACM_S2_WARP_BYTES (t.s2_hit_cap) : (size_t)t.bloom_words * 4 + warps * stage_bytes_per_warp;
  if (smem > img->smem_optin)
    return fail (ACM_B200_ERR_NOMEM, "filter tables do not fit shared memory%s", "");
  int rc;
  if ((rc = cx->d_cand_pos.ensure (p.cand_cap * 8)) || (rc = cx->d_cand_matches.ensure (p.cand_cap * 4)) || (rc = cx->d_cand_prefix.ensure (p.cand_cap * 4)) || (rc = cx->d_cand_inline.ensure (p.cand_cap * 16)) || (rc = cx->d_tile_first.ensure (p.ntiles * 8))
      || (rc = cx->d_tile_n.ensure (p.ntiles * 4)) || (rc = cx->d_tile_spill.ensure (p.ntiles * 4)) || (rc = cx->d_hot_spans.ensure ((size_t)p.hot_cap * 4 + 16)) || (rc = cx->d_counts.ensure (p.ntiles * 4)) || (rc = cx->d_offsets.ensure (p.ntiles * 8)))
    return rc;
  p.cand_pos = cx->d_cand_pos.as<uint64_t> ();
  p.cand_matches = cx->d_cand_matches.as<uint32_t> ();
  p.cand_prefix = cx->d_cand_prefix.as<uint32_t> ();
  p.cand_inline = cx->d_cand_inline.as<uint4> ();
  p.tile_first = cx->d_tile_first.as<uint64_t> ();
  p.tile_n = cx->d_tile_n.as<uint32_t> ();
  p.tile_spill = s2 ? cx->d_tile_spill.as<uint32_t> () : nullptr;
  p.hot_spans = cx->d_hot_spans.as<uint32_t> ();
  p.tile_matches = cx->d_counts.as<uint32_t> ();
  p.tile_offsets = cx->d_offsets.as<uint64_t> ();
  p.cand_count = &d_small->cand_count;
  p.overflow = &d_small->overflow;

  void (*f1) (const FilterParams) = nullptr;
  const bool ordered = dense; /* dense mode stages in position order (warp scan); the usual mode stages unordered and sorts the few survivors */
  const int K = 2; /* bits per key; 3 was measured slower, the kernels keep K as a template parameter */
#define ACM_F1_(Q_, K_, R_, O_) (p.bloom2 ? filter_scan_kernel<W, R_, Q_, K_, O_, true> : filter_scan_kernel<W, R_, Q_, K_, O_, false>)
#define ACM_F1(Q_, K_)                                                                                                                           \
  if (p.q == Q_ && K == K_)                                                                                                                      \
    f1 = ordered ? ACM_F1_ (Q_, K_, kDenseRows, true) : ACM_F1_ (Q_, K_, kRowsOpt, false)
  ACM_F1 (1, 2); ACM_F1 (2, 2);
  if (W == 1) {
    ACM_F1 (3, 2); ACM_F1 (4, 2);
  }
#undef ACM_F1
#undef ACM_F1_
  if (s2) /* 2 bits per key, kBatches x 32 hits confirmed at a time, 4 rows per tile: the measured optimum */
    f1 = m->option_s2_batches == 1 ? filter_scan_s2_kernel<2, 1, 4> : (m->option_s2_batches == 3 ? filter_scan_s2_kernel<2, 3, 4> : filter_scan_s2_kernel<2, 2, 4>);
  if (!f1)
    return fail (ACM_B200_ERR_INVALID, "no filter kernel for this window length%s", "");
  CUDA_TRY (cudaFuncSetAttribute (f1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cx->stats.smem_bytes = smem;
  cx->stats.filter_stride = s2 ? 2 : 1;
  /* header of Small (cand_count, grand_total, overflow, ...) cleared; the prefix symbols follow */
  cx->h_small->cand_count = 0;
  cx->h_small->grand_total = 0;
  cx->h_small->overflow = 0;
  cx->h_small->span_counter = 0;
  cx->h_small->hot_count = 0;
  CUDA_TRY (cudaMemcpyAsync (d_small, cx->h_small, offsetof (ScanContext::Small, prefix) + (size_t)job.prefix_len * 4, cudaMemcpyHostToDevice, job.st));
  const unsigned grid = (unsigned)std::min<uint64_t> (((s2 ? (job.n + 2047) / 2048 : p.ntiles) + warps - 1) / warps, (uint64_t)img->sm_count);
  if (first_segment)
    CUDA_TRY (cudaEventRecord (cx->ev[0], job.st));
  f1<<<grid, warps * 32, smem, job.st>>> (p);
  CUDA_TRY (cudaGetLastError ());
  if (first_segment)
    CUDA_TRY (cudaEventRecord (cx->ev[1], job.st));
  cx->stats.main_kernel_launches += 1;
  cx->stats.total_kernel_launches += 1;
  if (s2) { /* the spans whose stages overflowed are redone exactly; usually there are none and the kernel ends at once */
    CUDA_TRY (cudaFuncSetAttribute (filter_hot_spans_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHotSmemBytes));
    filter_hot_spans_kernel<<<(unsigned)std::min<uint64_t> (p.hot_cap, (uint64_t)img->sm_count * 2), 32, kHotSmemBytes, job.st>>> (p);
    CUDA_TRY (cudaGetLastError ());
    cx->stats.total_kernel_launches += 1;
  }
  const unsigned vgrid = (unsigned)std::min<uint64_t> ((p.cand_cap + 255) / 256, (uint64_t)img->sm_count * 8), tgrid = (unsigned)((p.ntiles + 255) / 256);
  filter_verify_kernel<W, false><<<vgrid, 256, 0, job.st>>> (p);
  CUDA_TRY (cudaGetLastError ());
  if (dense && !s2) /* tens of candidates per tile: one warp per tile */
    filter_tile_totals_dense_kernel<<<(unsigned)std::min<uint64_t> ((p.ntiles + 7) / 8, (uint64_t)img->sm_count * 64), 256, 0, job.st>>> (p);
  else
    filter_tile_totals_kernel<<<tgrid, 256, 0, job.st>>> (p);
  CUDA_TRY (cudaGetLastError ());
  cx->stats.total_kernel_launches += 2;
  if ((rc = device_exclusive_scan (cx, p.tile_matches, p.ntiles, cx->d_offsets.as<uint64_t> (), &d_small->grand_total, job.st)))
    return rc;
  /* the scalars the kernels wrote: overflow flags, number of candidates / unfinished spans, grand total */
  auto fetch_small = [&] () -> int {
    CUDA_TRY (cudaMemcpyAsync (cx->h_small, d_small, offsetof (ScanContext::Small, prefix), cudaMemcpyDeviceToHost, job.st));
    CUDA_TRY (cudaStreamSynchronize (job.st));
    if (cx->h_small->overflow || (s2 && cx->h_small->hot_count > p.hot_cap)) {
      if (dense)
        return fail (ACM_B200_ERR_CUDA, "candidate buffers overflowed in dense mode%s", "");
      return kFilterOverflow;
    }
    return ACM_B200_OK;
  };
  auto emit = [&] (uint64_t cap) -> int {
    p.matches = out;
    p.capacity = cap;
    filter_verify_kernel<W, true><<<vgrid, 256, 0, job.st>>> (p);
    CUDA_TRY (cudaGetLastError ());
    cx->stats.total_kernel_launches += 1;
    return ACM_B200_OK;
  };
  if (size_out_lazily) { /* single run into the library's own buffer: sized now that the total is known */
    if ((rc = fetch_small ()))
      return rc;
    const uint64_t want = std::min<uint64_t> (cx->h_small->grand_total, out_cap);
    if (last_segment)
      CUDA_TRY (cudaEventRecord (cx->ev[2], job.st));
    if (want) {
      if ((rc = cx->d_matches.ensure (want * sizeof (ACMB200Match))))
        return rc;
      out = cx->d_matches.as<ACMB200Match> ();
      if ((rc = emit (want)))
        return rc;
    }
    if (last_segment)
      CUDA_TRY (cudaEventRecord (cx->ev[3], job.st));
  } else { /* the record buffer is known: F4 writes what fits, the host looks at the scalars once everything is queued */
    if (last_segment)
      CUDA_TRY (cudaEventRecord (cx->ev[2], job.st));
    if (out_cap && out && (rc = emit (out_cap)))
      return rc;
    if (last_segment)
      CUDA_TRY (cudaEventRecord (cx->ev[3], job.st));
    if ((rc = fetch_small ()))
      return rc;
  }
  *total = cx->h_small->grand_total;
  cx->stats.last_nb_candidates += cx->h_small->cand_count;
  if (s2)
    cx->stats.hot_spans += cx->h_small->hot_count;
  *out_used = out;
  return ACM_B200_OK;
}

/* Filter engine: one run over the whole text in the sparse mode, with a candidate list sized from the candidate rate seen so far
 * (a first scan probes a prefix of the text for it: 1 Mi symbols, counted only); a list that overflows is retried once eight
 * times larger; a text dense in candidates -- by the probe, or because the retry overflowed too -- is scanned in the dense mode,
 * where no buffer can overflow and the text is processed in bounded segments (each re-reading max depth - 1 symbols of left
 * context) so that the scratch memory stays bounded. */
template <int W>
static int
run_filter (ACMachine *m, acm_device_image *img, ScanContext *cx, ScanJob &job, uint64_t *total, bool matches_on_device, ACMB200Match *user_matches) {
  const uint64_t kDenseSegment = 64ull << 20; /* symbols */
  const double kDenseRate = 1.0 / 64;         /* more candidates per symbol than this: the dense mode is cheaper */
  cx->stats.last_nb_candidates = 0;
  ACMB200Match *used = nullptr;
  int rc = kFilterOverflow;
  double rate = img->cand_rate;
  if (rate < 0 && job.n > (8u << 20)) { /* no history: look at a prefix first */
    ScanJob probe = job;
    probe.n = 1u << 20;
    probe.capacity = 0;
    uint64_t ignored = 0;
    rc = run_filter_once<W> (m, img, cx, probe, false, probe.n / 8, nullptr, 0, false, &ignored, &used, false, false);
    if (rc && rc != kFilterOverflow)
      return rc;
    rate = rc == kFilterOverflow ? 1.0 : (double)cx->h_small->cand_count / (double)probe.n;
    img->cand_rate = rate;
    cx->stats.last_nb_candidates = 0;
    rc = kFilterOverflow;
  }
  if (!img->prefer_dense && !(rate > kDenseRate)) {
    uint64_t cap = std::max<uint64_t> (1u << 18, (uint64_t)((double)job.n * std::max (1.0 / 512, 2 * rate)));
    for (int attempt = 0; attempt < 2 && rc == kFilterOverflow; attempt++, cap *= 8) {
      cx->stats.last_nb_candidates = 0;
      rc = run_filter_once<W> (m, img, cx, job, false, cap, user_matches, job.capacity, !matches_on_device, total, &used, true, true);
    }
    if (rc != kFilterOverflow) {
      if (!rc)
        img->cand_rate = (double)cx->stats.last_nb_candidates / (double)std::max<uint64_t> (job.n, 1);
      job.d_matches = used;
      return rc;
    }
    img->prefer_dense = true; /* texts this dense in candidates usually come in series: skip the doomed attempts next time */
    cx->stats.fallback_count++;
  }
  cx->stats.last_nb_candidates = 0;
  cx->stats.dense_scans++;
  ACMB200Match *out = user_matches;
  if (!matches_on_device && job.capacity) {
    if ((rc = cx->d_matches.ensure (job.capacity * sizeof (ACMB200Match))))
      return rc;
    out = cx->d_matches.as<ACMB200Match> ();
  }
  const uint64_t seg_lead = ((uint64_t)(m->max_depth ? m->max_depth - 1 : 0) + 15) / 16 * 16;
  uint64_t produced = 0, grand = 0;
  for (uint64_t start = 0; start < job.n; start += kDenseSegment) {
    ScanJob sub = job;
    const uint64_t lead = start ? seg_lead : 0, len = std::min<uint64_t> (kDenseSegment, job.n - start);
    sub.d_text = reinterpret_cast<const unsigned char *> (job.d_text) + (start - lead) * W;
    sub.n = len + lead;
    sub.lead = start ? std::max<uint64_t> (lead, job.lead > start - lead ? job.lead - (start - lead) : 0) : job.lead;
    sub.base = job.base + (start - lead);
    sub.prefix_len = start ? 0 : job.prefix_len;
    uint64_t seg_total = 0;
    const uint64_t room = job.capacity > produced ? job.capacity - produced : 0;
    rc = run_filter_once<W> (m, img, cx, sub, true, 0, out ? out + produced : nullptr, room, false, &seg_total, &used, start == 0, start + kDenseSegment >= job.n);
    if (rc)
      return rc;
    produced += std::min<uint64_t> (seg_total, room);
    grand += seg_total;
  }
  *total = grand;
  job.d_matches = out;
  if (cx->stats.last_nb_candidates < job.n / 64) { /* sparse again: the next scan may use the fast mode */
    img->prefer_dense = false;
    img->cand_rate = (double)cx->stats.last_nb_candidates / (double)std::max<uint64_t> (job.n, 1);
  }
  return ACM_B200_OK;
}

/* ---- cursor bookkeeping (host, at most max_depth symbols) ------------------------------------------------------------ */
static const ACState *
advance_cursor (ACMachine *m, const ACState *from, const unsigned char *tail, uint64_t tail_syms, bool tail_is_whole_text) {
  /* The state after the text is the longest suffix of (string(from) + text) that is a trie path; it is at most max_depth long, so
   * the last max_depth symbols decide it when the text is at least that long (walk from state 0), else walk everything from `from`. */
  const ACState *s = tail_is_whole_text ? from : m->root;
  const size_t w = acm_b200_symbol_width (m);
  const bool raw = m->symbol_kind == ACM_SYM_RAW1 || m->symbol_kind == ACM_SYM_RAW2 || m->symbol_kind == ACM_SYM_RAW4;
  std::vector<const void *> letter_of_class;
  if (!raw) {
    letter_of_class.assign ((size_t)m->nb_class + 1, nullptr);
    for (uint32_t k = 0; k < m->nb_class; k++)
      letter_of_class[m->class_sorted_id[k]] = m->class_letter[k];
  }
  for (uint64_t i = 0; i < tail_syms; i++) {
    const void *letter = tail + i * w;
    if (!raw) {
      uint32_t id;
      memcpy (&id, letter, 4);
      letter = id <= m->nb_class ? letter_of_class[id] : nullptr;
      if (!letter) { /* a letter of no keyword sends every state to state 0 */
        s = m->root;
        continue;
      }
    }
    s = acm_host_goto (s, letter);
  }
  return s;
}

/* ---- the scan entry point -------------------------------------------------------------------------------------------- */
/* counters of a finished scan into the machine's statistics (machine lock held) */
static void
merge_stats (ACMB200Stats &into, const ACMB200Stats &scan) {
  into.scan_kernel_ms = scan.scan_kernel_ms;
  into.main_kernel_ms = scan.main_kernel_ms;
  into.h2d_ms = scan.h2d_ms;
  into.d2h_ms = scan.d2h_ms;
  into.last_nb_symbols = scan.last_nb_symbols;
  into.last_nb_matches = scan.last_nb_matches;
  into.last_nb_candidates = scan.last_nb_candidates;
  into.smem_bytes = scan.smem_bytes;
  into.filter_stride = scan.filter_stride;
  into.main_kernel_launches += scan.main_kernel_launches;
  into.total_kernel_launches += scan.total_kernel_launches;
  into.fallback_count += scan.fallback_count;
  into.hot_spans += scan.hot_spans;
  into.dfa_event_scans += scan.dfa_event_scans;
  into.dense_scans += scan.dense_scans;
  into.dfa_tma_scans += scan.dfa_tma_scans;
  into.dfa_lean_scans += scan.dfa_lean_scans;
}

static int
scan_with_context (ACMachine *m, acm_device_image *img, ScanContext *cx, const ACMB200Scan *scan, uint64_t *nb_matches) {
  const acm_tables &t = img->tab;
  const size_t w = (size_t)t.width;
  const ACState *from = scan->cursor && *scan->cursor ? *scan->cursor : nullptr; /* null: state 0 */
  cudaStream_t st = scan->stream ? (cudaStream_t)scan->stream : cx->stream;
  uint64_t total = 0;
  int rc;
  ACMB200Stats &stats = cx->stats;
  stats = ACMB200Stats{};

  auto run_engine = [&] (ScanJob &job, uint64_t *tot, bool out_on_device, ACMB200Match *out) -> int {
    switch (t.engine) {
      case ACM_B200_ENGINE_DFA_SMEM:
        return run_dfa<uint16_t, true> (m, img, cx, job, tot, out_on_device, out);
      case ACM_B200_ENGINE_DFA_GLOBAL:
        return run_dfa<uint32_t, false> (m, img, cx, job, tot, out_on_device, out);
      default:
        return w == 1 ? run_filter<1> (m, img, cx, job, tot, out_on_device, out)
                      : (w == 2 ? run_filter<2> (m, img, cx, job, tot, out_on_device, out) : run_filter<4> (m, img, cx, job, tot, out_on_device, out));
    }
  };
  /* kernel times of the run that just finished (its events are recorded on st, already synchronised) */
  auto add_kernel_times = [&] () {
    float ms = 0, a = 0, b = 0;
    cudaEventElapsedTime (&ms, cx->ev[0], cx->ev[3]);
    stats.scan_kernel_ms += ms;
    cudaEventElapsedTime (&a, cx->ev[0], cx->ev[1]);
    if (t.engine != ACM_B200_ENGINE_FILTER)
      cudaEventElapsedTime (&b, cx->ev[2], cx->ev[3]);
    stats.main_kernel_ms += a + b;
  };
  /* carried cursor: DFA engines start chunk 0 from its state, the filter engine sees its string as a virtual prefix */
  auto apply_cursor = [&] (ScanJob &job) -> int {
    if (!from)
      return ACM_B200_OK;
    if (t.engine == ACM_B200_ENGINE_FILTER) {
      if (from->depth > 1024)
        return fail (ACM_B200_ERR_INVALID, "cursor deeper than 1024 symbols%s", "");
      job.prefix_len = from->depth;
      uint32_t k = from->depth;
      for (const ACState *s = from; s->parent; s = s->parent)
        cx->h_small->prefix[--k] = acm_symbol_of_state (m, s);
    } else
      job.init_dfa_state = t.dfa_of_state[from->id];
    return ACM_B200_OK;
  };
  const uint64_t stream_bytes = m->option_stream_bytes ? m->option_stream_bytes : (256ull << 20);

  if (scan->nb_symbols && (scan->text_on_device || scan->nb_symbols * w <= stream_bytes)) {
    /* ---- one run over the whole text ---- */
    ScanJob job = {};
    job.n = scan->nb_symbols;
    job.lead = scan->lead;
    job.base = scan->base;
    job.capacity = scan->capacity;
    job.st = st;
    if (scan->text_on_device) {
      if ((uintptr_t)scan->text & 15)
        return fail (ACM_B200_ERR_INVALID, "device text must be 16-byte aligned%s", "");
      job.d_text = scan->text;
    } else {
      if ((rc = cx->d_text.ensure (scan->nb_symbols * w + 64)))
        return rc;
      CUDA_TRY (cudaEventRecord (cx->ev[4], st));
      CUDA_TRY (cudaMemcpyAsync (cx->d_text.ptr, scan->text, scan->nb_symbols * w, cudaMemcpyHostToDevice, st));
      CUDA_TRY (cudaEventRecord (cx->ev[5], st));
      job.d_text = cx->d_text.ptr;
    }
    if ((rc = apply_cursor (job)) || (rc = run_engine (job, &total, scan->matches_on_device, scan->matches)))
      return rc;
    const uint64_t got = std::min<uint64_t> (total, scan->capacity);
    CUDA_TRY (cudaStreamSynchronize (st));
    if (!scan->text_on_device) {
      float ms = 0;
      cudaEventElapsedTime (&ms, cx->ev[4], cx->ev[5]);
      stats.h2d_ms = ms;
    }
    add_kernel_times ();
    if (got && !scan->matches_on_device) {
      const auto c0 = std::chrono::steady_clock::now ();
      CUDA_TRY (cudaMemcpyAsync (scan->matches, job.d_matches, got * sizeof (ACMB200Match), cudaMemcpyDeviceToHost, st));
      CUDA_TRY (cudaStreamSynchronize (st));
      stats.d2h_ms = std::chrono::duration<double, std::milli> (std::chrono::steady_clock::now () - c0).count ();
    }
  } else if (scan->nb_symbols) {
    /* ---- streaming ingest of host text: the text goes through two bounded device buffers; the copy of chunk
     * i+1 (copy stream) overlaps the scan of chunk i.  Every chunk after the first is copied with max depth - 1 symbols of left
     * context and scanned with that lead, exactly like a shard, so the records equal those of a single run.  The records of chunk
     * i leave on the copy stream too, from the record buffer of its parity, while chunk i+1 is scanned. ---- */
    const uint64_t chunk_syms = std::max<uint64_t> (stream_bytes / w / 4096 * 4096, 4096);
    const uint64_t seg_lead = ((uint64_t)(m->max_depth ? m->max_depth - 1 : 0) + 15) / 16 * 16;
    const unsigned char *host = reinterpret_cast<const unsigned char *> (scan->text);
    DevBuf *bufs[2] = { &cx->d_text, &cx->d_text2 };
    for (DevBuf *bf : bufs)
      if ((rc = bf->ensure ((chunk_syms + seg_lead) * w + 64)))
        return rc;
    const uint64_t nchunks = (scan->nb_symbols + chunk_syms - 1) / chunk_syms;
    auto copy_chunk = [&] (uint64_t i) -> cudaError_t {
      const uint64_t start = i * chunk_syms, lead = i ? seg_lead : 0, len = std::min<uint64_t> (chunk_syms, scan->nb_symbols - start);
      cudaError_t e = cudaMemcpyAsync (bufs[i & 1]->ptr, host + (start - lead) * w, (len + lead) * w, cudaMemcpyHostToDevice, cx->copy_stream);
      return e != cudaSuccess ? e : cudaEventRecord (cx->ev_copy[i & 1], cx->copy_stream);
    };
    const auto h0 = std::chrono::steady_clock::now ();
    CUDA_TRY (copy_chunk (0));
    uint64_t produced = 0;
    bool records_in_flight = false;
    for (uint64_t i = 0; i < nchunks; i++) {
      const uint64_t start = i * chunk_syms, lead = i ? seg_lead : 0, len = std::min<uint64_t> (chunk_syms, scan->nb_symbols - start);
      CUDA_TRY (cudaStreamWaitEvent (st, cx->ev_copy[i & 1], 0));
      if (i + 1 < nchunks) /* the other buffer was last read by the scan of chunk i-1, which has completed */
        CUDA_TRY (copy_chunk (i + 1));
      ScanJob job = {};
      job.d_text = bufs[i & 1]->ptr;
      job.n = len + lead;
      job.lead = i ? std::max<uint64_t> (lead, scan->lead > start - lead ? scan->lead - (start - lead) : 0) : scan->lead;
      job.base = scan->base + (start - lead);
      job.capacity = scan->capacity > produced ? scan->capacity - produced : 0;
      job.st = st;
      if (i == 0 && (rc = apply_cursor (job)))
        return rc;
      uint64_t seg_total = 0;
      /* records of this chunk go to the record buffer of its parity: the other one may still be on its way to the host */
      if (i & 1)
        std::swap (cx->d_matches, cx->d_matches2);
      rc = run_engine (job, &seg_total, false, nullptr);
      ACMB200Match *d_records = job.d_matches;
      if (i & 1)
        std::swap (cx->d_matches, cx->d_matches2);
      if (rc)
        return rc;
      CUDA_TRY (cudaStreamSynchronize (st));
      add_kernel_times ();
      const uint64_t got = std::min<uint64_t> (seg_total, job.capacity);
      if (got) { /* on the copy stream, behind the text copy of chunk i+1: the scan of chunk i+1 does not wait for it */
        CUDA_TRY (cudaMemcpyAsync (scan->matches + produced, d_records, got * sizeof (ACMB200Match), scan->matches_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                   cx->copy_stream));
        CUDA_TRY (cudaEventRecord (cx->ev_copy[2 + (i & 1)], cx->copy_stream));
        records_in_flight = true;
      }
      /* the record buffer of the NEXT chunk's parity must have left before that chunk's emit kernel writes it */
      if (i >= 1 && records_in_flight)
        CUDA_TRY (cudaStreamWaitEvent (st, cx->ev_copy[2 + ((i + 1) & 1)], 0));
      produced += got;
      total += seg_total;
    }
    CUDA_TRY (cudaStreamSynchronize (cx->copy_stream));
    stats.h2d_ms = std::chrono::duration<double, std::milli> (std::chrono::steady_clock::now () - h0).count (); /* whole pipeline, copies overlapped */
  }
  /* cursor out */
  if (scan->nb_symbols && scan->cursor) {
    const uint64_t tail = std::min<uint64_t> (scan->nb_symbols, m->max_depth);
    std::vector<unsigned char> buf (tail * w + 1);
    const unsigned char *src = reinterpret_cast<const unsigned char *> (scan->text) + (scan->nb_symbols - tail) * w;
    if (scan->text_on_device) {
      CUDA_TRY (cudaMemcpy (buf.data (), src, tail * w, cudaMemcpyDeviceToHost));
      src = buf.data ();
    }
    *scan->cursor = advance_cursor (m, from ? from : m->root, src, tail, scan->nb_symbols < m->max_depth);
  }
  stats.last_nb_symbols = scan->nb_symbols;
  stats.last_nb_matches = total;
  if (nb_matches)
    *nb_matches = total;
  return total > scan->capacity && scan->capacity ? ACM_B200_ERR_CAPACITY : ACM_B200_OK;
}

extern "C" int
acm_b200_scan_ex (ACMachine *m, const ACMB200Scan *scan, uint64_t *nb_matches) {
  if (!m || !scan || (!scan->text && scan->nb_symbols) || (scan->capacity && !scan->matches) || scan->lead > scan->nb_symbols)
    return fail (ACM_B200_ERR_INVALID, "invalid argument%s", "");
  DeviceGuard guard;
  if (scan->cursor) /* a carried cursor lives in the keyword trie */
    acm_ensure_trie (m);
  if (scan->cursor && *scan->cursor && (*scan->cursor)->machine != m)
    return fail (ACM_B200_ERR_INVALID, "the cursor belongs to another machine%s", "");
  /* machine lock: tables up to date, an image reference and a scan context of our own; then the GPU works without it */
  acm_lock (m);
  int rc = finalise_locked (m, -1);
  if (rc) {
    acm_unlock (m);
    return rc;
  }
  acm_device_image *img = m->device;
  ScanContext *cx = nullptr;
  for (ScanContext *c : img->contexts)
    if (!c->in_use) {
      cx = c;
      break;
    }
  if (!cx) {
    cx = new ScanContext ();
    if ((rc = cx->init ())) {
      cx->release ();
      delete cx;
      acm_unlock (m);
      return rc;
    }
    img->contexts.push_back (cx);
  }
  cx->in_use = true;
  img->refs++;
  acm_unlock (m);

  rc = scan_with_context (m, img, cx, scan, nb_matches);
  if (rc == ACM_B200_ERR_CUDA)
    cudaStreamSynchronize (cx->stream); /* leave nothing of a failed scan in flight on the context */

  acm_lock (m);
  cx->in_use = false;
  img->refs--;
  merge_stats (m->device->stats, cx->stats);
  if (img != m->device && img->refs == 0) { /* retired while we were scanning: we were its last user */
    for (acm_device_image *p = m->device; p; p = p->retired)
      if (p->retired == img) {
        p->retired = img->retired;
        img->retired = nullptr;
        cudaSetDevice (img->device);
        free_image (img);
        break;
      }
  }
  acm_unlock (m);
  return rc;
}

extern "C" int
acm_b200_scan (ACMachine *m, const ACState **cursor, const void *text, uint64_t nb_symbols, ACMB200Match *matches, uint64_t capacity, uint64_t *nb_matches) {
  ACMB200Scan s = {};
  s.text = text;
  s.nb_symbols = nb_symbols;
  s.matches = matches;
  s.capacity = capacity;
  s.cursor = cursor;
  return acm_b200_scan_ex (m, &s, nb_matches);
}
