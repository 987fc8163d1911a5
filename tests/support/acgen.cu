/* TEST / BENCH SUPPORT ONLY (not part of libac75.so): the position-addressable synthetic text of SURVEY.md 8(d).
 *
 *   raw64 (k) = splitmix64 finaliser (seed + k * 0x9E3779B97F4A7C15);  byte i = byte (i mod 8) of raw64 (i / 8)
 *   kind 0: bytes uniform 0..255;  kind 1: printable ASCII 0x20 + (b * 95 >> 8)
 *   plants: one keyword of the packed dictionary per `plant_period` bytes at a hashed offset; a plant may spill into the next
 *   period, where that period's own plant wins.
 * Any shard can be produced on the device (generate_text_kernel) and re-produced on the host (tests/support/textgen.py restates
 * the same generator in numpy; tests/test_textgen.py holds the two against each other).  Built into tests/support/libacgen.so. */
#include <cuda_runtime.h>
#include <stdint.h>

#define ACGEN_HD __host__ __device__ __forceinline__

struct GenParams {
  uint8_t *dst;
  uint64_t first, nb;
  int kind;
  uint64_t seed, plant_seed, plant_period;
  const uint8_t *dict_symbols;
  const uint64_t *dict_offsets;
  uint64_t dict_nb;
};

ACGEN_HD uint64_t
mix64 (uint64_t x) {
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

ACGEN_HD uint64_t
gen_raw64 (uint64_t seed, uint64_t k) {
  return mix64 (seed + k * 0x9E3779B97F4A7C15ull);
}

/* byte at absolute position i; identical on host and device */
ACGEN_HD uint8_t
gen_byte (const GenParams &g, uint64_t i) {
  if (g.plant_period && g.dict_nb) {
    const uint64_t blk = i / g.plant_period, off = i % g.plant_period;
    for (int back = 0; back < 2; back++) {
      if (back && blk == 0)
        break;
      const uint64_t b = blk - back;
      const uint64_t r = gen_raw64 (g.plant_seed, b);
      const uint64_t kw = (r >> 20) % g.dict_nb, at = (r & 0xFFFFFu) % g.plant_period;
      const uint64_t lo = g.dict_offsets[kw], len = g.dict_offsets[kw + 1] - lo;
      const uint64_t rel = off + (uint64_t)back * g.plant_period;
      if (rel >= at && rel < at + len)
        return g.dict_symbols[lo + (rel - at)];
    }
  }
  const uint8_t b = (uint8_t)(gen_raw64 (g.seed, i >> 3) >> (8 * (i & 7)));
  return g.kind == 1 ? (uint8_t)(0x20 + ((b * 95) >> 8)) : b;
}

__global__ void
generate_text_kernel (const __grid_constant__ GenParams g) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v * 16 < g.nb; v += stride) {
    const uint64_t at = v * 16;
    if (at + 16 <= g.nb) {
      uint32_t w[4] = { 0, 0, 0, 0 };
#pragma unroll
      for (int i = 0; i < 16; i++)
        w[i >> 2] |= (uint32_t)gen_byte (g, g.first + at + i) << (8 * (i & 3));
      *reinterpret_cast<uint4 *> (g.dst + at) = make_uint4 (w[0], w[1], w[2], w[3]);
    } else
      for (uint64_t i = at; i < g.nb; i++)
        g.dst[i] = gen_byte (g, g.first + i);
  }
}

/* 0 on success, a cudaError_t otherwise (-1: bad argument).  dst_on_device: dst is a 16-byte aligned device pointer on the
 * current device and the dictionary is copied there for the duration of the call; otherwise everything is host memory. */
extern "C" int
acgen_generate_text (void *dst, int dst_on_device, uint64_t first, uint64_t nb, int kind, uint64_t seed, uint64_t plant_seed, uint64_t plant_period,
                     const uint8_t *dict_symbols, const uint64_t *dict_offsets, uint64_t dict_nb, void *stream) {
  if (!dst && nb)
    return -1;
  GenParams g = {};
  g.dst = reinterpret_cast<uint8_t *> (dst);
  g.first = first;
  g.nb = nb;
  g.kind = kind;
  g.seed = seed;
  g.plant_seed = plant_seed;
  g.plant_period = plant_period;
  g.dict_nb = plant_period ? dict_nb : 0;
  if (!dst_on_device) {
    g.dict_symbols = dict_symbols;
    g.dict_offsets = dict_offsets;
    for (uint64_t i = 0; i < nb; i++)
      g.dst[i] = gen_byte (g, first + i);
    return 0;
  }
  if ((uintptr_t)dst & 15)
    return -1;
  cudaStream_t st = (cudaStream_t)stream;
  void *d_sym = nullptr, *d_off = nullptr;
  cudaError_t e = cudaSuccess;
  if (g.dict_nb) {
    const size_t sym_bytes = dict_offsets[dict_nb], off_bytes = (dict_nb + 1) * 8;
    if ((e = cudaMalloc (&d_sym, sym_bytes ? sym_bytes : 1)) == cudaSuccess && (e = cudaMalloc (&d_off, off_bytes)) == cudaSuccess
        && (e = cudaMemcpyAsync (d_sym, dict_symbols, sym_bytes, cudaMemcpyHostToDevice, st)) == cudaSuccess)
      e = cudaMemcpyAsync (d_off, dict_offsets, off_bytes, cudaMemcpyHostToDevice, st);
    g.dict_symbols = reinterpret_cast<const uint8_t *> (d_sym);
    g.dict_offsets = reinterpret_cast<const uint64_t *> (d_off);
  }
  if (e == cudaSuccess) {
    int sms = 148, dev = 0;
    cudaGetDevice (&dev); /* the current device, not device 0 */
    cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, dev);
    generate_text_kernel<<<sms * 8, 256, 0, st>>> (g);
    e = cudaGetLastError ();
    if (e == cudaSuccess)
      e = cudaStreamSynchronize (st);
  }
  if (d_sym)
    cudaFree (d_sym);
  if (d_off)
    cudaFree (d_off);
  return (int)e;
}
