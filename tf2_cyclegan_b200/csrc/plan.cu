// Host runtime: the Plan (a flat, pre-resolved list of kernel launches) and its
// Python bindings. Python compiles a network + batch size into a Plan ONCE —
// buffer addresses, TMA descriptors, tap tables and launch geometry are all
// baked into parameter blocks — and then `run()` replays it with no Python, no
// allocation and no descriptor encoding on the hot path; the whole range is
// also CUDA-graph capturable. This is the B200 counterpart of the reference's
// tf.function graph executor (main.py:197,269,325).
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "allreduce_adam.cuh"
#include "conv_gemm.cuh"
#include "elementwise.cuh"
#include "mx.cuh"
#include "wgrad_gemm.cuh"

#include <nvtx3/nvToolsExt.h>

namespace py = pybind11;
using namespace b2;

namespace b2 {
void set_kernel_options(int dbg, bool two_cta, int stages, int grid);   // conv_gemm.cu
}

namespace {

void check(cudaError_t e, const char* what) {
  if (e != cudaSuccess)
    throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    check(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q),
          "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)");
    if (ptr == nullptr || q != cudaDriverEntryPointSuccess)
      throw std::runtime_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// dtype: 0 = bf16, 1 = fp32, 2 = uint8 (fp8 payload). swizzle: 0 none, 3 = 128B.
py::bytes encode_tmap(uint64_t ptr, int dtype, std::vector<uint64_t> dims,
                      std::vector<uint64_t> strides_bytes, std::vector<uint32_t> box,
                      int swizzle) {
  const int rank = static_cast<int>(dims.size());
  if (rank < 1 || rank > 5 || static_cast<int>(box.size()) != rank ||
      static_cast<int>(strides_bytes.size()) != rank - 1)
    throw std::runtime_error("encode_tmap: bad rank / argument sizes");
  CUtensorMap m;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
  CUtensorMapDataType dt = dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                          : dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                       : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUtensorMapSwizzle sw = swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_128B
                         : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                         : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B
                                        : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = get_encode_fn()(&m, dt, rank, reinterpret_cast<void*>(ptr), gdim, gstr, bx, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw std::runtime_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(r));
  return py::bytes(reinterpret_cast<const char*>(&m), sizeof(m));
}

CUtensorMap tmap_from(const py::bytes& b) {
  std::string s = b;
  if (s.size() != sizeof(CUtensorMap)) throw std::runtime_error("tensor map must be 128 bytes");
  CUtensorMap m;
  memcpy(&m, s.data(), sizeof(m));
  return m;
}

using ViewT = std::tuple<uint64_t, long long, long long, long long>;
View view_from(const ViewT& t) {
  View v;
  v.p = reinterpret_cast<void*>(std::get<0>(t));
  v.sn = std::get<1>(t);
  v.sh = std::get<2>(t);
  v.sw = std::get<3>(t);
  return v;
}
template <class T>
T* P(uint64_t p) { return reinterpret_cast<T*>(p); }

using TapT = std::tuple<int, int, int, int, int>;   // map, dn, dw, dh, bk
void fill_taps(const std::vector<TapT>& taps, Tap* dst, int* n) {
  if (taps.size() > static_cast<size_t>(kMaxTaps)) throw std::runtime_error("too many taps");
  *n = static_cast<int>(taps.size());
  for (size_t i = 0; i < taps.size(); ++i) {
    dst[i].map = static_cast<int16_t>(std::get<0>(taps[i]));
    dst[i].dn = static_cast<int16_t>(std::get<1>(taps[i]));
    dst[i].dw = std::get<2>(taps[i]);
    dst[i].dh = std::get<3>(taps[i]);
    dst[i].bk = std::get<4>(taps[i]);
  }
}

class Plan {
 public:
  // Lanes: lane 0 runs on the caller's stream, lanes 1..4 on the plan's side streams (by convention
  // 0 / 1 = the two network pairs, 2 / 3 = their weight-gradient GEMMs, 4 = cross-GPU collectives in
  // ONE global order). Sync ops:
  //   kind 1 fork  : every side lane waits for lane 0
  //   kind 2 join  : lane 0 waits for every side lane that ran something since the last join
  //   kind 3 record: record event `ev` on lane `lane`
  //   kind 4 wait  : lane `lane` waits for the event recorded by op `ref`
  // All of it is plain event record / wait, so it also works under stream capture (CUDA graph).
  static constexpr int kLanes = 5;
  struct Op {
    std::string name;
    std::function<void(cudaStream_t)> fn;
    int lane = 0;
    int kind = 0;
    int ref = -1;
    cudaEvent_t ev = nullptr;
  };

  ~Plan() {
    for (auto& o : ops_)
      if (o.ev) cudaEventDestroy(o.ev);
    for (auto& s : side_)
      if (s) cudaStreamDestroy(s);
    for (auto& e : lane_ev_)
      if (e) cudaEventDestroy(e);
  }

  int add(std::string name, std::function<void(cudaStream_t)> fn) {
    Op o;
    o.name = std::move(name);
    o.fn = std::move(fn);
    o.lane = lane_;
    ops_.push_back(std::move(o));
    return static_cast<int>(ops_.size()) - 1;
  }
  void set_lane(int lane) {
    if (lane < 0 || lane >= kLanes) throw std::runtime_error("bad lane");
    lane_ = lane;
  }
  int add_sync(std::string name, int kind, int lane = 0, int ref = -1) {
    Op o;
    o.name = std::move(name);
    o.kind = kind;
    o.lane = lane;
    o.ref = ref;
    if (kind != 4) check(cudaEventCreateWithFlags(&o.ev, cudaEventDisableTiming), "cudaEventCreate");
    ops_.push_back(std::move(o));
    return static_cast<int>(ops_.size()) - 1;
  }
  int add_fork(std::string name) { return add_sync(std::move(name), 1); }
  int add_join(std::string name) { return add_sync(std::move(name), 2); }
  int add_event_record(std::string name, int lane) { return add_sync(std::move(name), 3, lane); }
  int add_event_wait(std::string name, int lane, int ref) {
    if (ref < 0 || ref >= size() || ops_[ref].kind != 3) throw std::runtime_error("bad event ref");
    return add_sync(std::move(name), 4, lane, ref);
  }
  int size() const { return static_cast<int>(ops_.size()); }
  void set_nvtx(bool on) { nvtx_ = on; }
  std::vector<std::string> names() const {
    std::vector<std::string> n;
    for (auto& o : ops_) n.push_back(o.name);
    return n;
  }

  cudaStream_t lane_stream(int lane, cudaStream_t main) {
    if (lane == 0) return main;
    if (side_[lane] == nullptr) {
      if (lane == kLanes - 1) {   // collectives: highest priority, so that their few CTAs are placed first
        int lo = 0, hi = 0;
        check(cudaDeviceGetStreamPriorityRange(&lo, &hi), "cudaDeviceGetStreamPriorityRange");
        check(cudaStreamCreateWithPriority(&side_[lane], cudaStreamNonBlocking, hi), "comm stream");
      } else {
        check(cudaStreamCreateWithFlags(&side_[lane], cudaStreamNonBlocking), "side stream");
      }
    }
    return side_[lane];
  }

  void join_dirty(cudaStream_t s, bool* dirty) {
    for (int l = 1; l < kLanes; ++l) {
      if (!dirty[l]) continue;
      if (lane_ev_[l] == nullptr)
        check(cudaEventCreateWithFlags(&lane_ev_[l], cudaEventDisableTiming), "cudaEventCreate");
      check(cudaEventRecord(lane_ev_[l], side_[l]), "join record");
      check(cudaStreamWaitEvent(s, lane_ev_[l], 0), "join wait");
      dirty[l] = false;
    }
  }

  void run(int begin, int end, uint64_t stream, bool check_each) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (end < 0 || end > size()) end = size();
    bool forked = false;
    bool dirty[kLanes] = {};
    for (int i = begin; i < end; ++i) {
      Op& o = ops_[i];
      if (o.kind != 0) {
        if (check_each) continue;          // debug mode: everything serial on the caller's stream
        if (o.kind == 1) {
          check(cudaEventRecord(o.ev, s), "fork record");
          for (int l = 1; l < kLanes; ++l)
            check(cudaStreamWaitEvent(lane_stream(l, s), o.ev, 0), "fork wait");
          forked = true;
        } else if (o.kind == 2) {
          join_dirty(s, dirty);
          forked = false;
        } else if (o.kind == 3) {
          if (forked || o.lane == 0) check(cudaEventRecord(o.ev, lane_stream(forked ? o.lane : 0, s)), "record");
        } else if (o.kind == 4) {
          // a wait whose record op lies outside [begin, end) (partial ranges) is skipped
          if (forked && o.ref >= begin)
            check(cudaStreamWaitEvent(lane_stream(o.lane, s), ops_[o.ref].ev, 0), "wait");
        }
        continue;
      }
      const int lane = forked ? o.lane : 0;
      if (lane != 0) dirty[lane] = true;
      if (nvtx_) nvtxRangePushA(o.name.c_str());
      o.fn(lane_stream(lane, s));
      if (nvtx_) nvtxRangePop();
      if (check_each) {
        cudaError_t e = cudaStreamSynchronize(s);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess)
          throw std::runtime_error("plan op #" + std::to_string(i) + " (" + ops_[i].name +
                                   ") failed: " + cudaGetErrorString(e));
      }
    }
    join_dirty(s, dirty);   // a range that ends inside a forked region always re-joins
    check(cudaGetLastError(), "plan launch");
  }

  // per-op device time (ms) measured with CUDA events; ops stay back-to-back on the stream
  std::vector<float> run_timed(int begin, int end, uint64_t stream) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (end < 0 || end > size()) end = size();
    const int n = end - begin;
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) check(cudaEventCreate(&e), "cudaEventCreate");
    check(cudaEventRecord(ev[0], s), "cudaEventRecord");
    for (int i = 0; i < n; ++i) {
      if (ops_[begin + i].kind == 0) ops_[begin + i].fn(s);
      check(cudaEventRecord(ev[i + 1], s), "cudaEventRecord");
    }
    check(cudaStreamSynchronize(s), "cudaStreamSynchronize");
    std::vector<float> ms(n);
    for (int i = 0; i < n; ++i) check(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]), "elapsed");
    for (auto& e : ev) cudaEventDestroy(e);
    return ms;
  }

  // ------------------------------------------------------------- GEMM ops
  int add_conv_gemm(std::string name, std::vector<py::bytes> a_maps, py::bytes b_map,
                    std::vector<TapT> taps, int cin_blocks, int tiles_w, int tiles_h, int bw,
                    int bh, int w_limit, uint64_t out, long long sn, long long sh, long long sw,
                    int cout, uint64_t bias, int act, uint64_t stats, int block_n, bool b_mn,
                    int num_images, std::vector<std::tuple<int, int, long long>> phases,
                    uint64_t out_scale, bool fp8, py::bytes ts_map, std::vector<int> ts, int h_limit) {
    auto p = std::make_shared<ConvGemmParams>();
    memset(p.get(), 0, sizeof(ConvGemmParams));
    if (a_maps.empty() || a_maps.size() > kMaxAMaps) throw std::runtime_error("bad a_maps");
    for (size_t i = 0; i < a_maps.size(); ++i) p->tmA[i] = tmap_from(a_maps[i]);
    p->tmB = tmap_from(b_map);
    fill_taps(taps, p->taps, &p->num_taps);
    if (bw * bh != 128) throw std::runtime_error("conv_gemm tile must be 128 pixels");
    p->cin_blocks = cin_blocks; p->tiles_w = tiles_w; p->tiles_h = tiles_h;
    p->bw = bw; p->bh = bh; p->w_limit = w_limit;
    p->h_limit = h_limit > 0 ? h_limit : tiles_h * bh;
    p->out = P<__nv_bfloat16>(out); p->out_sn = sn; p->out_sh = sh; p->out_sw = sw;
    p->cout = cout; p->bias = P<const float>(bias); p->act = act; p->stats = P<float>(stats);
    p->out_scale = P<const float>(out_scale); p->fp8 = fp8 ? 1 : 0;
    if (!ts.empty()) {   // tap sharing: {nbox, box_step, box_bytes, pitch, origin_dw, origin_dh, sbo}
      if (ts.size() != 7) throw std::runtime_error("ts needs 7 integers");
      p->tmTS = tmap_from(ts_map);
      p->ts = TapShare{1, ts[0], ts[1], ts[2], ts[3], ts[4], ts[5], ts[6]};
      for (int i = 0; i < p->num_taps; ++i) {
        const Tap& t = p->taps[i];
        const int shift = (t.dh - ts[5]) * ts[3] + (t.dw - ts[4]);
        if (t.map != 0 || t.dn != 0 || shift < 0 || shift * 128 + 15 * ts[6] + 1024 > ts[0] * ts[2])
          throw std::runtime_error("tap outside the shared halo tile");
      }
    }
    if (phases.empty()) phases.push_back({0, p->num_taps, 0});
    if (phases.size() > kMaxPhases) throw std::runtime_error("too many phases");
    p->num_phases = static_cast<int>(phases.size());
    for (size_t i = 0; i < phases.size(); ++i) {
      p->phase_tap_begin[i] = std::get<0>(phases[i]);
      p->phase_tap_count[i] = std::get<1>(phases[i]);
      p->phase_out_off[i] = std::get<2>(phases[i]);
    }
    return add(std::move(name), [p, block_n, b_mn, num_images](cudaStream_t s) {
      launch_conv_gemm(*p, block_n, b_mn, num_images, s);
    });
  }

  // MXFP8 conv (conv_gemm_mx.cu): E4M3 activations / weights with E8M0 scales per 32 K elements
  int add_conv_gemm_mx(std::string name, py::bytes a_map, py::bytes sfa_map, py::bytes b_map, uint64_t sfb,
                       std::vector<TapT> taps, int cin_blocks, int tiles_w, int tiles_h, int bw, int bh,
                       int w_limit, int h_limit, uint64_t out, long long sn, long long sh, long long sw,
                       int cout, uint64_t bias, int act, uint64_t stats, int num_images, py::bytes ts_map,
                       py::bytes sfts_map, std::vector<int> ts) {
    auto p = std::make_shared<ConvGemmMxParams>();
    memset(p.get(), 0, sizeof(ConvGemmMxParams));
    ConvGemmParams& g = p->g;
    g.tmA[0] = tmap_from(a_map);
    g.tmB = tmap_from(b_map);
    p->tmSFA = tmap_from(sfa_map);
    p->sfb = P<const uint8_t>(sfb);
    fill_taps(taps, g.taps, &g.num_taps);
    if (bw * bh != 128) throw std::runtime_error("conv_gemm_mx tile must be 128 pixels");
    g.cin_blocks = cin_blocks; g.tiles_w = tiles_w; g.tiles_h = tiles_h;
    g.bw = bw; g.bh = bh; g.w_limit = w_limit;
    g.h_limit = h_limit > 0 ? h_limit : tiles_h * bh;
    g.out = P<__nv_bfloat16>(out); g.out_sn = sn; g.out_sh = sh; g.out_sw = sw;
    g.cout = cout; g.bias = P<const float>(bias); g.act = act; g.stats = P<float>(stats);
    g.fp8 = 1; g.num_phases = 1; g.phase_tap_count[0] = g.num_taps;
    if (!ts.empty()) {   // tap sharing, as in add_conv_gemm; the scale records get a halo box of the same pixels
      if (ts.size() != 7) throw std::runtime_error("ts needs 7 integers");
      g.tmTS = tmap_from(ts_map);
      p->tmSFTS = tmap_from(sfts_map);
      g.ts = TapShare{1, ts[0], ts[1], ts[2], ts[3], ts[4], ts[5], ts[6]};
      for (int i = 0; i < g.num_taps; ++i) {
        const Tap& t = g.taps[i];
        const int shift = (t.dh - ts[5]) * ts[3] + (t.dw - ts[4]);
        if (t.map != 0 || t.dn != 0 || shift < 0 || shift * 128 + 15 * ts[6] + 1024 > ts[0] * ts[2])
          throw std::runtime_error("tap outside the shared halo tile");
      }
    }
    return add(std::move(name), [p, num_images](cudaStream_t s) { launch_conv_gemm_mx(*p, num_images, s); });
  }
  int add_mx_weight_quant(std::string name, uint64_t w, uint64_t w8, uint64_t sfb, int rows, int taps, int kdim,
                          long long s_row, long long s_tap, long long s_k) {
    MxWeightQuantParams p{P<const __nv_bfloat16>(w), P<uint8_t>(w8), P<uint8_t>(sfb), rows, taps, kdim,
                          s_row, s_tap, s_k};
    return add(std::move(name), [p](cudaStream_t s) { launch_mx_weight_quant(p, s); });
  }

  int add_wgrad(std::string name, std::vector<py::bytes> dy_maps, std::vector<py::bytes> a_maps,
                std::vector<TapT> taps, int ktiles_w, int ktiles_h, int kbw, int kbh,
                int num_images, int cin, int cout, uint64_t dw, long long ldw, int block_n,
                int ksplit) {
    auto p = std::make_shared<WgradParams>();
    memset(p.get(), 0, sizeof(WgradParams));
    if (dy_maps.empty() || dy_maps.size() > kMaxAMaps) throw std::runtime_error("bad dy_maps");
    for (size_t i = 0; i < dy_maps.size(); ++i) p->tmDY[i] = tmap_from(dy_maps[i]);
    if (a_maps.empty() || a_maps.size() > kMaxAMaps) throw std::runtime_error("bad a_maps");
    for (size_t i = 0; i < a_maps.size(); ++i) p->tmA[i] = tmap_from(a_maps[i]);
    fill_taps(taps, p->taps, &p->num_taps);
    if (kbw * kbh != 64) throw std::runtime_error("wgrad K tile must be 64 pixels");
    p->ktiles_w = ktiles_w; p->ktiles_h = ktiles_h; p->kbw = kbw; p->kbh = kbh;
    p->num_images = num_images; p->cin = cin; p->cout = cout;
    p->dw = P<float>(dw); p->ldw = ldw;
    return add(std::move(name), [p, block_n, ksplit](cudaStream_t s) {
      launch_wgrad_gemm(*p, block_n, ksplit, s);
    });
  }

  // ------------------------------------------------------ elementwise ops
  int add_in_apply(std::string name, ViewT x, uint64_t stats, uint64_t gamma, uint64_t beta,
                   float inv_count, float eps, int act, ViewT skip, ViewT dst, int halo, int N,
                   int H, int W, int C, ViewT dst8, uint64_t scale8, uint64_t amax8, ViewT sf8) {
    InApplyParams p{view_from(x), P<const float>(stats), P<const float>(gamma),
                    P<const float>(beta), inv_count, eps, act, view_from(skip), view_from(dst),
                    halo, N, H, W, C, view_from(dst8), P<const float>(scale8), P<float>(amax8),
                    view_from(sf8)};
    return add(std::move(name), [p](cudaStream_t s) { launch_in_apply(p, s); });
  }
  int add_in_bwd_reduce(std::string name, ViewT g1, int fold1, ViewT g2, ViewT x, uint64_t stats,
                        uint64_t gamma, uint64_t beta, float inv_count, float eps, int act,
                        ViewT dy, uint64_t bstats, int N, int H, int W, int C) {
    InBwdReduceParams p{view_from(g1), fold1, view_from(g2), view_from(x),
                        P<const float>(stats), P<const float>(gamma), P<const float>(beta),
                        inv_count, eps, act, view_from(dy), P<float>(bstats), N, H, W, C};
    return add(std::move(name), [p](cudaStream_t s) { launch_in_bwd_reduce(p, s); });
  }
  int add_in_bwd_apply(std::string name, ViewT dy, ViewT x, uint64_t stats, uint64_t bstats,
                       uint64_t gamma, float inv_count, float eps, ViewT dx, uint64_t dgamma,
                       uint64_t dbeta, int N, int H, int W, int C, ViewT dx8, ViewT sf8) {
    InBwdApplyParams p{view_from(dy), view_from(x), P<const float>(stats),
                       P<const float>(bstats), P<const float>(gamma), inv_count, eps,
                       view_from(dx), P<float>(dgamma), P<float>(dbeta), N, H, W, C,
                       view_from(dx8), view_from(sf8)};
    return add(std::move(name), [p](cudaStream_t s) { launch_in_bwd_apply(p, s); });
  }
  int add_act_bwd(std::string name, ViewT g1, int fold1, ViewT g2, ViewT out, int act, ViewT dy,
                  uint64_t dbias, int c_valid, int N, int H, int W, int C) {
    ActBwdParams p{view_from(g1), fold1, view_from(g2), view_from(out), act, view_from(dy),
                   P<float>(dbias), c_valid, N, H, W, C};
    return add(std::move(name), [p](cudaStream_t s) { launch_act_bwd(p, s); });
  }
  int add_img_prep(std::string name, uint64_t src_f32, ViewT src_view, ViewT dst_reflect,
                   int reflect_pad, ViewT dst_plain, int N, int H, int W) {
    ImgPrepParams p{P<const void>(src_f32), src_f32 != 0 ? 1 : 0, view_from(src_view),
                    view_from(dst_reflect), reflect_pad, view_from(dst_plain), N, H, W};
    return add(std::move(name), [p](cudaStream_t s) { launch_img_prep(p, s); });
  }
  int add_l1_loss(std::string name, ViewT pred, uint64_t target, float loss_scale,
                  float grad_scale, uint64_t loss_slot, uint64_t err_slot, float err_scale,
                  ViewT g, int N, int H, int W) {
    L1LossParams p{view_from(pred), P<const float>(target), loss_scale, grad_scale,
                   P<float>(loss_slot), P<float>(err_slot), err_scale, view_from(g), N, H, W};
    return add(std::move(name), [p](cudaStream_t s) { launch_l1_loss(p, s); });
  }
  int add_mse_loss(std::string name, ViewT d, float target, float loss_scale, float grad_scale,
                   uint64_t loss_slot, ViewT g, int N, int H, int W) {
    MseLossParams p{view_from(d), target, loss_scale, grad_scale, P<float>(loss_slot),
                    view_from(g), N, H, W};
    return add(std::move(name), [p](cudaStream_t s) { launch_mse_loss(p, s); });
  }
  // kind: 0 = weight amax, 1 = scales (from amaxes), 2 = weight quantisation
  int add_fp8_op(std::string name, int kind, std::vector<long long> w_off, std::vector<long long> w_numel,
                 uint64_t w, uint64_t w8, uint64_t act_amax, uint64_t w_amax, uint64_t act_scale,
                 uint64_t w_scale, uint64_t out_scale, float default_act_scale) {
    auto t = std::make_shared<Fp8Table>();
    memset(t.get(), 0, sizeof(Fp8Table));
    if (w_off.size() > kMaxFp8Layers || w_off.size() != w_numel.size()) throw std::runtime_error("fp8 table");
    t->num_layers = static_cast<int>(w_off.size());
    for (size_t i = 0; i < w_off.size(); ++i) { t->w_off[i] = w_off[i]; t->w_numel[i] = w_numel[i]; }
    t->w = P<const __nv_bfloat16>(w); t->w8 = P<uint8_t>(w8); t->act_amax = P<float>(act_amax);
    t->w_amax = P<float>(w_amax); t->act_scale = P<float>(act_scale); t->w_scale = P<float>(w_scale);
    t->out_scale = P<float>(out_scale); t->default_act_scale = default_act_scale;
    return add(std::move(name), [t, kind](cudaStream_t s) {
      if (kind == 0) launch_fp8_w_amax(*t, s);
      else if (kind == 1) launch_fp8_scales(*t, s);
      else launch_fp8_w_quant(*t, s);
    });
  }
  int add_tap_gather(std::string name, uint64_t q, uint64_t bias, ViewT out, int act, int k, int Hp,
                     int Wp, int N, int H, int W) {
    TapGatherParams p{P<const __nv_bfloat16>(q), P<const float>(bias), view_from(out), act, k, Hp, Wp,
                      N, H, W};
    return add(std::move(name), [p](cudaStream_t s) { launch_tap_gather(p, s); });
  }
  int add_adam(std::string name, uint64_t w, uint64_t m, uint64_t v, uint64_t g, uint64_t w_bf16,
               uint64_t step, float lr, float b1, float b2, float eps, float grad_scale,
               long long n) {
    AdamParams p{P<float>(w), P<float>(m), P<float>(v), P<const float>(g),
                 P<__nv_bfloat16>(w_bf16), P<const int>(step), lr, b1, b2, eps, grad_scale, n};
    return add(std::move(name), [p](cudaStream_t s) { launch_adam(p, s); });
  }
  int add_step_increment(std::string name, uint64_t step) {
    return add(std::move(name), [step](cudaStream_t s) { launch_step_increment(P<int>(step), s); });
  }
  int add_mask_grad(std::string name, uint64_t g, uint64_t mask, long long n) {
    return add(std::move(name),
               [g, mask, n](cudaStream_t s) { launch_mask_grad(P<float>(g), P<const uint8_t>(mask), n, s); });
  }
  // dst[i] = bf16(src[idx[i]]) (0 where idx < 0); src is fp32, or bf16 when src_is_bf16
  int add_gather_bf16(std::string name, uint64_t dst, uint64_t src, uint64_t idx, long long n,
                      bool src_is_bf16) {
    return add(std::move(name), [dst, src, idx, n, src_is_bf16](cudaStream_t s) {
      if (src_is_bf16)
        launch_gather_bf16_from_bf16(P<__nv_bfloat16>(dst), P<const __nv_bfloat16>(src), P<const int>(idx), n, s);
      else
        launch_gather_bf16(P<__nv_bfloat16>(dst), P<const float>(src), P<const int>(idx), n, s);
    });
  }
  int add_scatter_add(std::string name, uint64_t dst, uint64_t src, uint64_t idx, long long n) {
    return add(std::move(name), [dst, src, idx, n](cudaStream_t s) {
      launch_scatter_add(P<float>(dst), P<const float>(src), P<const int>(idx), n, s);
    });
  }
  int add_memset(std::string name, uint64_t ptr, long long nbytes) {
    return add(std::move(name), [ptr, nbytes](cudaStream_t s) {
      check(cudaMemsetAsync(P<void>(ptr), 0, nbytes, s), "cudaMemsetAsync");
    });
  }
  int add_d4_bwd(std::string name, ViewT gd, ViewT a, uint64_t w, ViewT dx, uint64_t dw,
                 uint64_t dbias, int pad_before, int N, int H, int W, int C, bool dgrad,
                 bool wgrad) {
    D4BwdParams p{view_from(gd), view_from(a), P<const __nv_bfloat16>(w), view_from(dx),
                  P<float>(dw), P<float>(dbias), pad_before, N, H, W, C};
    return add(std::move(name), [p, dgrad, wgrad](cudaStream_t s) {
      if (dgrad) launch_d4_dgrad(p, s);
      if (wgrad) launch_d4_wgrad(p, s);
    });
  }
  // grad / wbf16 / w / m / v are BUFFER BASES; [lo, lo + n) is the two-shot (sharded) region, [v_lo, v_lo + v_n)
  // the one-shot (replicated) region, sc_* the loss-scalar side channel (see allreduce_adam.cuh)
  int add_allreduce_adam(std::string name, std::vector<uint64_t> grad_ptrs,
                         std::vector<uint64_t> wbf16_ptrs, std::vector<uint64_t> signal_ptrs,
                         uint64_t w, uint64_t m, uint64_t v, uint64_t step, float lr, float b1,
                         float b2, float eps, long long lo, long long n, long long v_lo, long long v_n,
                         std::vector<uint64_t> sc_src, uint64_t sc_dst, int sc_n, int rank, int world,
                         int num_blocks, uint64_t mc_grad, uint64_t mc_wbf16,
                         std::vector<uint64_t> stage_ptrs) {
    auto p = std::make_shared<AllreduceAdamParams>();
    memset(p.get(), 0, sizeof(AllreduceAdamParams));
    if (world > kMaxRanks || world < 1) throw std::runtime_error("bad world size");
    if (static_cast<int>(grad_ptrs.size()) != world || static_cast<int>(wbf16_ptrs.size()) != world ||
        static_cast<int>(signal_ptrs.size()) != world)
      throw std::runtime_error("allreduce_adam: need one pointer per rank");
    if ((lo | n | v_lo | v_n) % 8 != 0 || n < 0 || v_n < 0) throw std::runtime_error("allreduce_adam: regions must be multiples of 8");
    if (sc_n < 0 || sc_n > kMaxScalars || (sc_n > 0 && static_cast<int>(sc_src.size()) != world))
      throw std::runtime_error("allreduce_adam: bad scalar side channel");
    if (num_blocks < 1 || num_blocks > 148) throw std::runtime_error("allreduce_adam: 1..148 blocks");
    if (!stage_ptrs.empty() && static_cast<int>(stage_ptrs.size()) != world)
      throw std::runtime_error("allreduce_adam: push mode needs one staging pointer per rank");
    for (int r = 0; r < world; ++r) {
      p->grad[r] = P<float>(grad_ptrs[r]);
      p->w_bf16[r] = P<__nv_bfloat16>(wbf16_ptrs[r]);
      p->signal[r] = P<uint32_t>(signal_ptrs[r]);
      if (sc_n > 0) p->sc_src[r] = P<const float>(sc_src[r]);
      if (!stage_ptrs.empty()) p->stage[r] = P<float>(stage_ptrs[r]);
    }
    p->w = P<float>(w); p->m = P<float>(m); p->v = P<float>(v); p->step = P<const int>(step);
    p->lr = lr; p->beta1 = b1; p->beta2 = b2; p->eps = eps;
    p->lo = lo; p->n = n; p->v_lo = v_lo; p->v_n = v_n;
    p->sc_dst = P<float>(sc_dst); p->sc_n = sc_n; p->rank = rank;
    p->world = world; p->mc_grad = P<float>(mc_grad); p->mc_w_bf16 = P<__nv_bfloat16>(mc_wbf16);
    return add(std::move(name), [p, num_blocks](cudaStream_t s) {
      launch_allreduce_adam(*p, num_blocks, s);
    });
  }

 private:
  std::vector<Op> ops_;
  int lane_ = 0;
  // NVTX range per op (name = the op label) for timeline tools; on by B2_NVTX=1 or set_nvtx()
  bool nvtx_ = false;
  cudaStream_t side_[kLanes] = {};
  cudaEvent_t lane_ev_[kLanes] = {};
};

}  // namespace

PYBIND11_MODULE(_C, m) {
  m.doc() = "b200-cyclegan native runtime (sm_100a kernels + plan executor)";
  m.def("encode_tmap", &encode_tmap, py::arg("ptr"), py::arg("dtype"), py::arg("dims"),
        py::arg("strides_bytes"), py::arg("box"), py::arg("swizzle") = 3);
  m.def("set_kernel_options", &b2::set_kernel_options, py::arg("kernel_debug"), py::arg("two_cta"),
        py::arg("stages"), py::arg("grid"), "process-wide kernel options (tf2_cyclegan_b200/config.py)");
  m.def("set_ew_stream", &b2::set_ew_stream, "InstanceNorm kernels: true = bulk-copy ring (default), false = direct");
  m.def("ew_stream", &b2::ew_stream);
  m.def("set_mx_profile", [](uint64_t ptr) { b2::set_mx_profile(reinterpret_cast<unsigned long long*>(ptr)); },
        "device buffer of 16 uint64 counters for the MXFP8 kernel's stall attribution (0 = off)");
  m.def("device_sync", []() { check(cudaDeviceSynchronize(), "cudaDeviceSynchronize"); });
  m.attr("MAX_TAPS") = kMaxTaps;
  py::class_<Plan>(m, "Plan")
      .def(py::init<>())
      .def("size", &Plan::size)
      .def("names", &Plan::names)
      .def("set_nvtx", &Plan::set_nvtx)
      .def("run", &Plan::run, py::arg("begin") = 0, py::arg("end") = -1, py::arg("stream") = 0,
           py::arg("check_each") = false, py::call_guard<py::gil_scoped_release>())
      .def("set_lane", &Plan::set_lane)
      .def("add_fork", &Plan::add_fork)
      .def("add_join", &Plan::add_join)
      .def("add_event_record", &Plan::add_event_record)
      .def("add_event_wait", &Plan::add_event_wait)
      .def("run_timed", &Plan::run_timed, py::arg("begin") = 0, py::arg("end") = -1,
           py::arg("stream") = 0)
      .def("add_conv_gemm", &Plan::add_conv_gemm, py::arg("name"), py::arg("a_maps"), py::arg("b_map"),
           py::arg("taps"), py::arg("cin_blocks"), py::arg("tiles_w"), py::arg("tiles_h"), py::arg("bw"),
           py::arg("bh"), py::arg("w_limit"), py::arg("out"), py::arg("sn"), py::arg("sh"), py::arg("sw"),
           py::arg("cout"), py::arg("bias"), py::arg("act"), py::arg("stats"), py::arg("block_n"),
           py::arg("b_mn"), py::arg("num_images"), py::arg("phases"), py::arg("out_scale") = 0,
           py::arg("fp8") = false, py::arg("ts_map") = py::bytes(), py::arg("ts") = std::vector<int>(),
           py::arg("h_limit") = 0)
      .def("add_wgrad", &Plan::add_wgrad)
      .def("add_in_apply", &Plan::add_in_apply, py::arg("name"), py::arg("x"), py::arg("stats"),
           py::arg("gamma"), py::arg("beta"), py::arg("inv_count"), py::arg("eps"), py::arg("act"),
           py::arg("skip"), py::arg("dst"), py::arg("halo"), py::arg("N"), py::arg("H"), py::arg("W"),
           py::arg("C"), py::arg("dst8") = ViewT{0, 0, 0, 0}, py::arg("scale8") = 0, py::arg("amax8") = 0,
           py::arg("sf8") = ViewT{0, 0, 0, 0})
      .def("add_conv_gemm_mx", &Plan::add_conv_gemm_mx, py::arg("name"), py::arg("a_map"), py::arg("sfa_map"),
           py::arg("b_map"), py::arg("sfb"), py::arg("taps"), py::arg("cin_blocks"), py::arg("tiles_w"),
           py::arg("tiles_h"), py::arg("bw"), py::arg("bh"), py::arg("w_limit"), py::arg("h_limit"),
           py::arg("out"), py::arg("sn"), py::arg("sh"), py::arg("sw"), py::arg("cout"), py::arg("bias"),
           py::arg("act"), py::arg("stats"), py::arg("num_images"), py::arg("ts_map") = py::bytes(""),
           py::arg("sfts_map") = py::bytes(""), py::arg("ts") = std::vector<int>{})
      .def("add_mx_weight_quant", &Plan::add_mx_weight_quant)
      .def("add_fp8_op", &Plan::add_fp8_op)
      .def("add_in_bwd_reduce", &Plan::add_in_bwd_reduce)
      .def("add_in_bwd_apply", &Plan::add_in_bwd_apply, py::arg("name"), py::arg("dy"), py::arg("x"), py::arg("stats"),
           py::arg("bstats"), py::arg("gamma"), py::arg("inv_count"), py::arg("eps"), py::arg("dx"),
           py::arg("dgamma"), py::arg("dbeta"), py::arg("N"), py::arg("H"), py::arg("W"), py::arg("C"),
           py::arg("dx8") = ViewT{0, 0, 0, 0}, py::arg("sf8") = ViewT{0, 0, 0, 0})
      .def("add_act_bwd", &Plan::add_act_bwd)
      .def("add_img_prep", &Plan::add_img_prep)
      .def("add_l1_loss", &Plan::add_l1_loss)
      .def("add_mse_loss", &Plan::add_mse_loss)
      .def("add_tap_gather", &Plan::add_tap_gather)
      .def("add_adam", &Plan::add_adam)
      .def("add_step_increment", &Plan::add_step_increment)
      .def("add_mask_grad", &Plan::add_mask_grad)
      .def("add_gather_bf16", &Plan::add_gather_bf16, py::arg("name"), py::arg("dst"), py::arg("src"),
           py::arg("idx"), py::arg("n"), py::arg("src_is_bf16") = false)
      .def("add_scatter_add", &Plan::add_scatter_add)
      .def("add_memset", &Plan::add_memset)
      .def("add_d4_bwd", &Plan::add_d4_bwd)
      .def("add_allreduce_adam", &Plan::add_allreduce_adam, py::arg("name"), py::arg("grad_ptrs"),
           py::arg("wbf16_ptrs"), py::arg("signal_ptrs"), py::arg("w"), py::arg("m"), py::arg("v"), py::arg("step"),
           py::arg("lr"), py::arg("b1"), py::arg("b2"), py::arg("eps"), py::arg("lo"), py::arg("n"),
           py::arg("v_lo"), py::arg("v_n"), py::arg("sc_src"), py::arg("sc_dst"), py::arg("sc_n"), py::arg("rank"),
           py::arg("world"), py::arg("num_blocks"), py::arg("mc_grad"), py::arg("mc_wbf16"),
           py::arg("stage_ptrs") = std::vector<uint64_t>());
}
