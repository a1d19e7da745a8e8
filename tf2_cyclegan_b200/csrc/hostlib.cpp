// Host-side native helpers (plain C ABI, loaded with ctypes; no CUDA, no Python headers):
//   * CRC-32C (Castagnoli) with the SSE4.2 instruction — TFRecord event files and TensorBundle
//     checkpoints checksum every record / tensor with it;
//   * the input pipeline's per-image map, fused: random flip -> bilinear resize (half-pixel
//     centres, no antialias = tf.image.resize) -> crop -> /127.5 - 1, computing only the cropped
//     window, over a batch of images on a small thread pool (reference main.py:35-50 runs these
//     as tf.data C++ kernels with num_parallel_calls=AUTOTUNE);
//   * the batch gather into a pinned staging buffer (tf.data's batch()).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#if defined(__SSE4_2__)
#include <nmmintrin.h>
#endif

namespace {

uint32_t g_table[8][256];
bool g_table_ready = false;

void init_table() {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
    g_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t)
      g_table[t][i] = (g_table[t - 1][i] >> 8) ^ g_table[0][g_table[t - 1][i] & 0xFF];
  g_table_ready = true;
}

template <typename F>
void parallel_for(int n, int threads, F&& fn) {
  threads = std::max(1, std::min(threads, n));
  if (threads == 1) {
    for (int i = 0; i < n; ++i) fn(i);
    return;
  }
  std::vector<std::thread> pool;
  pool.reserve(threads);
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&, t] {
      for (int i = t; i < n; i += threads) fn(i);
    });
  for (auto& th : pool) th.join();
}

struct Axis {   // one output coordinate of the bilinear resize
  int lo, hi;
  float w;
};

// tf.image.resize(bilinear, half_pixel_centers, antialias=False) restricted to [first, first+n)
void make_axis(int in_size, int out_size, int first, int n, bool flip, std::vector<Axis>& ax) {
  ax.resize(n);
  const float scale = static_cast<float>(in_size) / static_cast<float>(out_size);
  for (int i = 0; i < n; ++i) {
    const float src = (static_cast<float>(first + i) + 0.5f) * scale - 0.5f;
    const float fl = std::floor(src);
    int lo = std::max(static_cast<int>(fl), 0);
    int hi = std::min(static_cast<int>(std::ceil(src)), in_size - 1);
    lo = std::min(lo, in_size - 1);
    if (flip) { lo = in_size - 1 - lo; hi = in_size - 1 - hi; }
    ax[i] = {lo, hi, src - fl};
  }
}

}  // namespace

extern "C" {

uint32_t b2h_crc32c(const uint8_t* p, size_t n, uint32_t crc) {
  crc = ~crc;
#if defined(__SSE4_2__)
  while (n && (reinterpret_cast<uintptr_t>(p) & 7)) { crc = _mm_crc32_u8(crc, *p++); --n; }
  uint64_t c = crc;
  while (n >= 8) {
    uint64_t v;
    std::memcpy(&v, p, 8);
    c = _mm_crc32_u64(c, v);
    p += 8; n -= 8;
  }
  crc = static_cast<uint32_t>(c);
  while (n--) crc = _mm_crc32_u8(crc, *p++);
#else
  if (!g_table_ready) init_table();
  while (n >= 8) {
    uint64_t v;
    std::memcpy(&v, p, 8);
    v ^= crc;
    crc = g_table[7][v & 0xFF] ^ g_table[6][(v >> 8) & 0xFF] ^ g_table[5][(v >> 16) & 0xFF] ^
          g_table[4][(v >> 24) & 0xFF] ^ g_table[3][(v >> 32) & 0xFF] ^
          g_table[2][(v >> 40) & 0xFF] ^ g_table[1][(v >> 48) & 0xFF] ^ g_table[0][v >> 56];
    p += 8; n -= 8;
  }
  while (n--) crc = g_table[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8);
#endif
  return ~crc;
}

// One image: uint8 HWC (C channels) -> float32 [crop_h, crop_w, C] = normalise(crop(resize(flip(src)))).
void b2h_preprocess_image(const uint8_t* src, int H, int W, int C, float* dst, int flip,
                          int resize_h, int resize_w, int top, int left, int crop_h, int crop_w) {
  std::vector<Axis> ay, ax;
  make_axis(H, resize_h, top, crop_h, false, ay);
  make_axis(W, resize_w, left, crop_w, flip != 0, ax);
  // separable: horizontally resized source rows are cached (consecutive outputs share them),
  // the vertical blend + normalisation is a plain vectorisable loop
  const int row_len = crop_w * C;
  std::vector<float> rows(2 * static_cast<size_t>(row_len));
  int cached[2] = {-1, -1};
  auto hrow = [&](int sy, int slot) {
    if (cached[slot] == sy) return;
    const uint8_t* r = src + static_cast<size_t>(sy) * W * C;
    float* o = rows.data() + static_cast<size_t>(slot) * row_len;
    for (int x = 0; x < crop_w; ++x) {
      const uint8_t* p0 = r + ax[x].lo * C;
      const uint8_t* p1 = r + ax[x].hi * C;
      const float wx = ax[x].w;
      for (int c = 0; c < C; ++c) {
        const float v0 = p0[c];
        o[x * C + c] = v0 + (static_cast<float>(p1[c]) - v0) * wx;
      }
    }
    cached[slot] = sy;
  };
  const float k = 1.0f / 127.5f;
  for (int y = 0; y < crop_h; ++y) {
    int s0 = 0, s1 = 1;
    if (cached[1] == ay[y].lo || cached[0] == ay[y].hi) { s0 = 1; s1 = 0; }
    hrow(ay[y].lo, s0);
    if (ay[y].hi == ay[y].lo) s1 = s0; else hrow(ay[y].hi, s1);
    const float* t0 = rows.data() + static_cast<size_t>(s0) * row_len;
    const float* t1 = rows.data() + static_cast<size_t>(s1) * row_len;
    const float wy = ay[y].w;
    float* out = dst + static_cast<size_t>(y) * row_len;
    for (int i = 0; i < row_len; ++i) out[i] = (t0[i] + (t1[i] - t0[i]) * wy) * k - 1.0f;
  }
}

// A batch of same-sized images; params[i] = {flip, top, left}.
void b2h_preprocess_batch(const uint8_t* src, int n, int H, int W, int C, float* dst,
                          const int32_t* params, int resize_h, int resize_w, int crop_h,
                          int crop_w, int threads) {
  const size_t in_sz = static_cast<size_t>(H) * W * C;
  const size_t out_sz = static_cast<size_t>(crop_h) * crop_w * C;
  parallel_for(n, threads, [&](int i) {
    b2h_preprocess_image(src + i * in_sz, H, W, C, dst + i * out_sz, params[3 * i],
                         resize_h, resize_w, params[3 * i + 1], params[3 * i + 2], crop_h, crop_w);
  });
}

// dst[i] = src[idx[i]] for rows of row_bytes bytes.
void b2h_gather_rows(uint8_t* dst, const uint8_t* src, const int64_t* idx, int n,
                     size_t row_bytes, int threads) {
  parallel_for(n, threads, [&](int i) {
    std::memcpy(dst + static_cast<size_t>(i) * row_bytes,
                src + static_cast<size_t>(idx[i]) * row_bytes, row_bytes);
  });
}

}  // extern "C"
