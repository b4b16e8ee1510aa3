// saber_b200 — integer / indexing stages of automatic mask generation, bit-exact given identical
// low-res logits: bilinear up-sampling of 256x256 mask logits to the crop size fused with the
// stability score (two thresholded popcounts), binarisation, bounding box, near-crop-edge test
// and bit-packing into the full image frame (warp ballots; the fp32 full-res logits that upstream
// materialises per mask — 805 MB per 64-point batch at 1024^2 — never exist), plus greedy box NMS
// with torchvision semantics. Restates sam2/utils/amg.py (calculate_stability_score,
// batched_mask_to_box, is_box_near_crop_edge, uncrop_masks), SAM2Transforms.postprocess_masks and
// torchvision.ops.nms as driven by sam2/automatic_mask_generator.py::_process_batch/_process_crop
// (SURVEY §8a U5, Appendix A2). The oracle twin is oracle/amg_post_ref.py.
#include "common.cuh"

namespace {

// ATen area_pixel_compute_source_index (align_corners=False, not cubic). The expression tree — including
// where the fused multiply-adds sit — is the one torch's CPU F.interpolate(bilinear) evaluates (pinned
// bitwise in tests/test_oracle_pins.py), so the thresholded masks are bit-identical to the reference's CPU path.
__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1,
                                          float& l0, float& l1) {
  float s = __fmaf_rn(scale, __fadd_rn(static_cast<float>(dst), 0.5f), -0.5f);
  if (s < 0.f) s = 0.f;
  i0 = min(static_cast<int>(floorf(s)), in_size - 1);
  l1 = fminf(fmaxf(__fsub_rn(s, static_cast<float>(i0)), 0.f), 1.f);
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l0 = __fsub_rn(1.f, l1);
}

struct MaskPostParams {
  const float* planes;     // low-res logit planes [B, 4, S, S] (all four mask tokens per prompt)
  const float* ious4;      // [B, 4] predicted IoU per token
  const int* sel;          // [n] chosen token per prompt (single-mask / m2m mode), or null
  int cpp;                 // candidates per prompt: 1 (sel given, or token 0) or 3 (multimask tokens 1..3)
  int n, S;
  int Hc, Wc, x0, y0, H, W, WW;  // crop size / origin, full frame size, words per row
  float pred_iou_thresh, mask_thresh, stab_offset, stab_thresh, edge_atol;
  int x1, y1;              // crop box far corner (exclusive)
  unsigned char* keep;     // [n]
  float* stability;        // [n]
  float* iou_out;          // [n]
  int* bbox;               // [n,4] full-frame xyxy (inclusive max), upstream uncrop_boxes_xyxy applied
  int* area;               // [n]
  uint32_t* bits;          // [n, H, WW]
  const int* geom;         // optional device-side {Hc, Wc, x0, y0, base}: overrides the crop geometry and offsets the
                           // outputs by `base` slots (lets one captured CUDA graph serve every crop / batch)
};

// 1024 threads: one warp per 32-column strip of a 1024-wide frame. The row walk of a warp is a serial chain of gathers,
// so a candidate's time is (rows walked per warp) x (latency of one row): 8 warps per candidate left an SM with 8-16
// resident warps and took 1.29 ms per 192 candidates; 32 warps quarter the chain (profiles/r02zzf, r02zzg).
__global__ void __launch_bounds__(1024)
mask_post_kernel(MaskPostParams p) {
  if (p.geom) {
    p.Hc = p.geom[0];
    p.Wc = p.geom[1];
    p.x0 = p.geom[2];
    p.y0 = p.geom[3];
    p.x1 = p.x0 + p.Wc;
    p.y1 = p.y0 + p.Hc;
    const long long base = p.geom[4];
    p.keep += base;
    p.stability += base;
    p.iou_out += base;
    p.bbox += 4 * base;
    p.area += base;
    p.bits += base * p.H * p.WW;
  }
  const int i = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int prompt = i / p.cpp;
  const int token = p.sel ? p.sel[prompt] : (p.cpp == 3 ? 1 + i % 3 : 0);
  const int plane_id = prompt * 4 + token;
  const float iou = p.ious4[plane_id];
  uint32_t* bits = p.bits + static_cast<long long>(i) * p.H * p.WW;
  const bool iou_ok = !(p.pred_iou_thresh > 0.f) || (iou > p.pred_iou_thresh);
  if (tid == 0) p.iou_out[i] = iou;
  if (!iou_ok) {
    if (tid == 0) {
      p.keep[i] = 0;
      p.stability[i] = 0.f;
      p.area[i] = 0;
      p.bbox[4 * i + 0] = p.bbox[4 * i + 1] = p.bbox[4 * i + 2] = p.bbox[4 * i + 3] = 0;
    }
    return;  // block-uniform
  }
  const float* plane = p.planes + static_cast<long long>(plane_id) * p.S * p.S;
  const float scale_h = __fdiv_rn(static_cast<float>(p.S), static_cast<float>(p.Hc));
  const float scale_w = __fdiv_rn(static_cast<float>(p.S), static_cast<float>(p.Wc));
  const float hi = __fadd_rn(p.mask_thresh, p.stab_offset), lo = __fsub_rn(p.mask_thresh, p.stab_offset);
  const bool same = (p.Hc == p.S && p.Wc == p.S);

  // Column-word-major: a warp owns 32-pixel column strips of the frame and walks down the rows, so the x half of the
  // source-index arithmetic (and the crop test in x) is hoisted out of the row loop, the y half is warp-uniform, and
  // the statistics are per-lane counters (one ballot per word instead of three + popcounts). The packed words of 32
  // consecutive rows are collected one per lane and stored together. The per-pixel expression tree is unchanged
  // (bit-exact masks); the row-major version spent ~2 M warp instructions per up-sampled candidate.
  int inter = 0, uni = 0, area = 0;
  int minx = 1 << 30, maxx = -1, miny = 1 << 30, maxy = -1;  // crop-frame coordinates
  const int nwarps = blockDim.x >> 5;
  for (int wx = warp; wx < p.WW; wx += nwarps) {
    const int x = wx * 32 + lane;
    const int cx = x - p.x0;
    const bool col_any = (wx * 32 + 31 >= p.x0) && (wx * 32 < p.x0 + p.Wc);  // warp-uniform
    const bool in_x = (cx >= 0 && cx < p.Wc && x < p.W);
    int x0i = 0, x1i = 0;
    float lx0 = 0.f, lx1 = 0.f;
    if (in_x && !same) src_index(scale_w, cx, p.S, x0i, x1i, lx0, lx1);
    bool lane_any = false;  // this lane's column has a mask pixel
    uint32_t mine = 0;      // packed word of row (yb + lane) of the current 32-row group
#pragma unroll 4
    for (int y = 0; y < p.H; ++y) {
      const int cy = y - p.y0;
      uint32_t word = 0;
      if (col_any && cy >= 0 && cy < p.Hc) {  // warp-uniform
        float val = -INFINITY;
        if (in_x) {
          if (same) {
            val = __ldg(plane + cy * p.S + cx);
          } else {
            int y0i, y1i;
            float ly0, ly1;
            src_index(scale_h, cy, p.S, y0i, y1i, ly0, ly1);
            // read-only path: lets the compiler hoist the gathers of the unrolled rows above the packed-word stores
            const float v00 = __ldg(plane + y0i * p.S + x0i), v01 = __ldg(plane + y0i * p.S + x1i);
            const float v10 = __ldg(plane + y1i * p.S + x0i), v11 = __ldg(plane + y1i * p.S + x1i);
            const float t0 = __fmaf_rn(v00, lx0, __fmul_rn(v01, lx1));
            const float t1 = __fmaf_rn(v10, lx0, __fmul_rn(v11, lx1));
            val = __fmaf_rn(t0, ly0, __fmul_rn(t1, ly1));
          }
        }
        const bool on = in_x && val > p.mask_thresh;
        inter += (in_x && val > hi) ? 1 : 0;
        uni += (in_x && val > lo) ? 1 : 0;
        word = __ballot_sync(0xffffffffu, on);
        if (on) {
          ++area;
          lane_any = true;
          miny = min(miny, cy);
          maxy = max(maxy, cy);
        }
      }
      if ((y & 31) == lane) mine = word;
      if ((y & 31) == 31 || y == p.H - 1) {  // flush up to 32 rows: lane l holds the word of row (y & ~31) + l
        const int yy = (y & ~31) + lane;
        if (yy <= y) bits[yy * p.WW + wx] = mine;
        mine = 0;
      }
    }
    if (lane_any) {
      minx = min(minx, cx);
      maxx = max(maxx, cx);
    }
  }
  // per-lane statistics -> lane 0 of every warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    inter += __shfl_xor_sync(0xffffffffu, inter, o);
    uni += __shfl_xor_sync(0xffffffffu, uni, o);
    area += __shfl_xor_sync(0xffffffffu, area, o);
    minx = min(minx, __shfl_xor_sync(0xffffffffu, minx, o));
    maxx = max(maxx, __shfl_xor_sync(0xffffffffu, maxx, o));
    miny = min(miny, __shfl_xor_sync(0xffffffffu, miny, o));
    maxy = max(maxy, __shfl_xor_sync(0xffffffffu, maxy, o));
  }
  __shared__ int s_red[32][7];
  if (lane == 0) {
    s_red[warp][0] = inter;
    s_red[warp][1] = uni;
    s_red[warp][2] = area;
    s_red[warp][3] = minx;
    s_red[warp][4] = maxx;
    s_red[warp][5] = miny;
    s_red[warp][6] = maxy;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < nwarps; ++w) {
      inter += s_red[w][0];
      uni += s_red[w][1];
      area += s_red[w][2];
      minx = min(minx, s_red[w][3]);
      maxx = max(maxx, s_red[w][4]);
      miny = min(miny, s_red[w][5]);
      maxy = max(maxy, s_red[w][6]);
    }
    // torch: int32 / int32 -> fp32 true division (0/0 = NaN, and NaN >= thr is false)
    const float stab = __fdiv_rn(static_cast<float>(inter), static_cast<float>(uni));
    bool keep = true;
    if (p.stab_thresh > 0.f) keep = stab >= p.stab_thresh;
    int bx0 = 0, by0 = 0, bx1 = 0, by1 = 0;  // empty mask -> [0,0,0,0] (batched_mask_to_box)
    if (area > 0) {
      bx0 = minx;
      by0 = miny;
      bx1 = maxx;
      by1 = maxy;
    }
    // is_box_near_crop_edge on un-cropped boxes: near the crop box but not near the image box
    const float fb[4] = {static_cast<float>(bx0 + p.x0), static_cast<float>(by0 + p.y0),
                         static_cast<float>(bx1 + p.x0), static_cast<float>(by1 + p.y0)};
    const float cb[4] = {static_cast<float>(p.x0), static_cast<float>(p.y0), static_cast<float>(p.x1),
                         static_cast<float>(p.y1)};
    const float ob[4] = {0.f, 0.f, static_cast<float>(p.W), static_cast<float>(p.H)};
    bool near = false;
    for (int k = 0; k < 4; ++k) {
      const bool nc = fabsf(fb[k] - cb[k]) <= p.edge_atol;
      const bool ni = fabsf(fb[k] - ob[k]) <= p.edge_atol;
      near = near || (nc && !ni);
    }
    keep = keep && !near;
    p.keep[i] = keep ? 1 : 0;
    p.stability[i] = stab;
    p.area[i] = area;
    p.bbox[4 * i + 0] = bx0 + p.x0;
    p.bbox[4 * i + 1] = by0 + p.y0;
    p.bbox[4 * i + 2] = bx1 + p.x0;
    p.bbox[4 * i + 3] = by1 + p.y0;
  }
}

// ---------------------------------------------------------------------------------------------
// NMS (torchvision.ops.nms semantics): stable sort by score descending, suppress iff IoU > thr.
// Device-resident: candidate lists and their lengths live in device memory, so a whole image's AMG
// (21 crops + the cross-crop pass) runs without a host synchronisation.
// ---------------------------------------------------------------------------------------------

// cand[0..*count) = { base + i : keep[base + i] != 0, i < n } in ascending order (single block).
__global__ void __launch_bounds__(1024)
compact_keep_kernel(const unsigned char* __restrict__ keep, int base, int n, int* __restrict__ cand,
                    int* __restrict__ count) {
  __shared__ int warp_tot[32];
  __shared__ int running;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    const bool k = i < n && keep[base + i] != 0;
    const uint32_t bal = __ballot_sync(0xffffffffu, k);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = running;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (k) cand[off + __popc(bal & ((1u << lane) - 1u))] = base + i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 32; ++w) t += warp_tot[w];
      running += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = running;
}

// order[rank] = cand[i], rank = #{j : score_j > score_i or (score_j == score_i and j < i)} (stable, desc).
__global__ void __launch_bounds__(256)
nms_rank_kernel(const float* __restrict__ scores, const int* __restrict__ cand, const int* __restrict__ n_ptr,
                int* __restrict__ order) {
  __shared__ float tile[256];
  const int n = *n_ptr;
  if (blockIdx.x * 256 >= n) return;  // block-uniform
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float si = i < n ? scores[cand[i]] : 0.f;
  int rank = 0;
  for (int j0 = 0; j0 < n; j0 += 256) {
    const int j = j0 + threadIdx.x;
    tile[threadIdx.x] = j < n ? scores[cand[j]] : 0.f;
    __syncthreads();
    const int lim = min(256, n - j0);
    if (i < n) {
      for (int t = 0; t < lim; ++t) {
        const float sj = tile[t];
        rank += (sj > si) || (sj == si && (j0 + t) < i);
      }
    }
    __syncthreads();
  }
  if (i < n) order[rank] = cand[i];
}

__device__ __forceinline__ float4 box_f(const int* __restrict__ bbox, int slot) {
  const int4 b = *reinterpret_cast<const int4*>(bbox + 4 * static_cast<long long>(slot));
  return make_float4(static_cast<float>(b.x), static_cast<float>(b.y), static_cast<float>(b.z),
                     static_cast<float>(b.w));
}

__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, float thr) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float w = fmaxf(__fsub_rn(right, left), 0.f), h = fmaxf(__fsub_rn(bottom, top), 0.f);
  const float inter = __fmul_rn(w, h);
  const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float sb_ = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb_), inter)) > thr;
}

// mask[r, cw] bit c = IoU(sorted box r, sorted box cw*64 + c) > thr for c-index > r
__global__ void __launch_bounds__(64)
nms_mask_kernel(const int* __restrict__ bbox, const int* __restrict__ order, const int* __restrict__ n_ptr,
                float thr, unsigned long long* __restrict__ mask, int col_blocks_cap) {
  const int n = *n_ptr;
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (cb < rb || cb * 64 >= n || rb * 64 >= n) return;  // only the upper triangle is ever read
  __shared__ float4 cbox[64];
  const int csize = min(64, n - cb * 64), rsize = min(64, n - rb * 64);
  if (threadIdx.x < csize) cbox[threadIdx.x] = box_f(bbox, order[cb * 64 + threadIdx.x]);
  __syncthreads();
  if (threadIdx.x < rsize) {
    const int r = rb * 64 + threadIdx.x;
    const float4 a = box_f(bbox, order[r]);
    unsigned long long t = 0;
    const int start = (rb == cb) ? threadIdx.x + 1 : 0;
    for (int c = start; c < csize; ++c)
      if (iou_gt(a, cbox[c], thr)) t |= 1ULL << c;
    mask[static_cast<long long>(r) * col_blocks_cap + cb] = t;
  }
}

// Sequential greedy pass, 64 boxes at a time: one thread resolves the chunk's diagonal block, then
// all threads OR the kept rows into the running suppression bitmap. Kept slots are appended (in
// score-descending order) to out_list at *out_count, which is then advanced.
__global__ void __launch_bounds__(256)
nms_scan_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ order,
                const int* __restrict__ n_ptr, int col_blocks_cap, int* __restrict__ out_list,
                int* __restrict__ out_count) {
  extern __shared__ unsigned long long remv[];  // col_blocks words
  __shared__ unsigned long long kept_bits;
  __shared__ int nkeep;
  const int n = *n_ptr;
  const int col_blocks = (n + 63) / 64;
  const int out_base = *out_count;
  for (int w = threadIdx.x; w < col_blocks; w += blockDim.x) remv[w] = 0;
  if (threadIdx.x == 0) nkeep = 0;
  __syncthreads();
  for (int cb = 0; cb < col_blocks; ++cb) {
    if (threadIdx.x == 0) {
      unsigned long long r = remv[cb], kept = 0;
      const int csize = min(64, n - cb * 64);
      for (int c = 0; c < csize; ++c) {
        if (!((r >> c) & 1ULL)) {
          kept |= 1ULL << c;
          r |= mask[static_cast<long long>(cb * 64 + c) * col_blocks_cap + cb];
          out_list[out_base + nkeep++] = order[cb * 64 + c];
        }
      }
      kept_bits = kept;
    }
    __syncthreads();
    const unsigned long long kept = kept_bits;
    for (int w = cb + 1 + threadIdx.x; w < col_blocks; w += blockDim.x) {
      unsigned long long acc = remv[w], kb = kept;
      while (kb) {
        const int c = __ffsll(static_cast<long long>(kb)) - 1;
        kb &= kb - 1;
        acc |= mask[static_cast<long long>(cb * 64 + c) * col_blocks_cap + w];
      }
      remv[w] = acc;
    }
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x == 0) *out_count = out_base + nkeep;
}

// inter[i, j] (i < j) = popcount(mask_i & mask_j) when min(area)/max(area) >= area_ratio_thresh (fp64, as
// REF saber/segmenters/utils.py:31-45 evaluates it) else -1. Only the rows / words inside both bounding
// boxes are scanned (bits outside a mask's box are zero). One warp per pair.
__global__ void __launch_bounds__(256)
pair_inter_kernel(const uint32_t* __restrict__ bits, const int* __restrict__ bbox, const int* __restrict__ area,
                  int m, int H, int WW, double area_ratio_thresh, int* __restrict__ inter) {
  const int lane = threadIdx.x & 31;
  const long long npairs = static_cast<long long>(m) * m;
  const long long nwords = static_cast<long long>(H) * WW;
  for (long long pi = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; pi < npairs;
       pi += (static_cast<long long>(gridDim.x) * blockDim.x) >> 5) {
    const int i = static_cast<int>(pi / m), j = static_cast<int>(pi % m);
    if (j <= i) continue;
    const int ai = area[i], aj = area[j];
    const int amax = max(ai, aj), amin = min(ai, aj);
    const double ratio = amax > 0 ? static_cast<double>(amin) / static_cast<double>(amax) : 0.0;
    int result = -1;
    if (!(ratio < area_ratio_thresh)) {
      const int x0 = max(bbox[4 * i + 0], bbox[4 * j + 0]), y0 = max(bbox[4 * i + 1], bbox[4 * j + 1]);
      const int x1 = min(bbox[4 * i + 2], bbox[4 * j + 2]), y1 = min(bbox[4 * i + 3], bbox[4 * j + 3]);
      int cnt = 0;
      if (x1 >= x0 && y1 >= y0) {
        const int w0 = x0 >> 5, w1 = x1 >> 5, nw = w1 - w0 + 1;
        const int tot = (y1 - y0 + 1) * nw;
        const uint32_t* bi = bits + static_cast<long long>(i) * nwords;
        const uint32_t* bj = bits + static_cast<long long>(j) * nwords;
        for (int t = lane; t < tot; t += 32) {
          const long long o = static_cast<long long>(y0 + t / nw) * WW + w0 + t % nw;
          cnt += __popc(bi[o] & bj[o]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      result = cnt;
    }
    if (lane == 0) inter[pi] = result;
  }
}

// rows[sel[k]] of a packed [n, H, WW] bit volume -> bool bytes [m, H, W]
__global__ void __launch_bounds__(256)
unpack_bits_kernel(const uint32_t* __restrict__ bits, const int* __restrict__ sel, int m, int H, int W,
                   int WW, unsigned char* __restrict__ out) {
  const long long total = static_cast<long long>(m) * H * W;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const long long t = i / W;
    const int y = static_cast<int>(t % H);
    const int k = static_cast<int>(t / H);
    const int src = sel ? sel[k] : k;
    const uint32_t w = bits[(static_cast<long long>(src) * H + y) * WW + (x >> 5)];
    out[i] = (w >> (x & 31)) & 1u;
  }
}

// dst[k] = src[sel[k]] for rows of `row_words` 32-bit words (compaction of kept masks / records)
__global__ void __launch_bounds__(256)
gather_rows_kernel(const uint32_t* __restrict__ src, const int* __restrict__ sel, int m,
                   long long row_words, uint32_t* __restrict__ dst) {
  const long long total = static_cast<long long>(m) * row_words;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long k = i / row_words, o = i - k * row_words;
    dst[i] = src[static_cast<long long>(sel[k]) * row_words + o];
  }
}

}  // namespace

extern "C" int sb_amg_mask_post(const float* planes, const float* ious4, const int* sel, int cpp,
                                int n, int S, int Hc, int Wc, int x0, int y0,
                                int H, int W, float pred_iou_thresh, float mask_thresh,
                                float stab_offset, float stab_thresh, unsigned char* keep,
                                float* stability, float* iou_out, int* bbox, int* area, void* bits,
                                const int* geom_dev, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(geom_dev != nullptr || (n > 0 && S > 0 && Hc > 0 && Wc > 0 && H >= y0 + Hc && W >= x0 + Wc),
             "sb_amg_mask_post: bad geometry n=%d S=%d crop=%dx%d+%d+%d frame=%dx%d", n, S, Hc, Wc,
             x0, y0, H, W);
  MaskPostParams p;
  SB_REQUIRE(cpp == 1 || (cpp == 3 && sel == nullptr), "sb_amg_mask_post: cpp must be 1, or 3 without sel");
  p.planes = planes;
  p.ious4 = ious4;
  p.sel = sel;
  p.cpp = cpp;
  p.n = n;
  p.S = S;
  p.Hc = Hc;
  p.Wc = Wc;
  p.x0 = x0;
  p.y0 = y0;
  p.H = H;
  p.W = W;
  p.WW = (W + 31) / 32;
  p.pred_iou_thresh = pred_iou_thresh;
  p.mask_thresh = mask_thresh;
  p.stab_offset = stab_offset;
  p.stab_thresh = stab_thresh;
  p.edge_atol = 20.0f;
  p.x1 = x0 + Wc;
  p.y1 = y0 + Hc;
  p.keep = keep;
  p.stability = stability;
  p.iou_out = iou_out;
  p.bbox = bbox;
  p.area = area;
  p.bits = static_cast<uint32_t*>(bits);
  p.geom = geom_dev;
  mask_post_kernel<<<n, 1024, 0, stream>>>(p);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// cand[0..count) <- slots base..base+n with keep != 0 (ascending); count is a device int.
extern "C" int sb_compact_keep(const unsigned char* keep, int base, int n, int* cand, int* count,
                               void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(n >= 0 && base >= 0, "sb_compact_keep: bad range");
  compact_keep_kernel<<<1, 1024, 0, stream>>>(keep, base, n, cand, count);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// Greedy NMS over the candidate slots cand[0..*n_ptr) (n_cap = host-side upper bound of *n_ptr).
// bbox [*,4] int32 xyxy and scores [*] fp32 are indexed by slot. Workspace: order [n_cap] int32,
// mask_ws [n_cap * ceil(n_cap/64)] u64. Kept slots are appended in score-descending order to
// out_list[*out_count ...] and *out_count advanced (device ints; no host synchronisation).
extern "C" int sb_nms_dev(const int* bbox, const float* scores, const int* cand, const int* n_ptr, int n_cap,
                          float iou_thresh, int* order, void* mask_ws, int* out_list, int* out_count,
                          void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(n_cap > 0, "sb_nms_dev: empty capacity");
  SB_REQUIRE((reinterpret_cast<uintptr_t>(bbox) & 15) == 0, "sb_nms_dev: bbox must be 16-byte aligned");
  const int col_blocks = (n_cap + 63) / 64;
  SB_REQUIRE(col_blocks * 8 <= 200 * 1024, "sb_nms_dev: too many boxes (%d)", n_cap);
  nms_rank_kernel<<<(n_cap + 255) / 256, 256, 0, stream>>>(scores, cand, n_ptr, order);
  SB_CHECK_LAUNCH();
  dim3 grid(col_blocks, col_blocks);
  nms_mask_kernel<<<grid, 64, 0, stream>>>(bbox, order, n_ptr, iou_thresh,
                                           static_cast<unsigned long long*>(mask_ws), col_blocks);
  SB_CHECK_LAUNCH();
  const size_t smem = static_cast<size_t>(col_blocks) * 8;
  if (smem > 48 * 1024)
    SB_CHECK_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
  nms_scan_kernel<<<1, 256, smem, stream>>>(static_cast<const unsigned long long*>(mask_ws), order, n_ptr,
                                            col_blocks, out_list, out_count);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// Pairwise mask intersections of m packed masks [m, H, ceil(W/32)] (rows i < j of inter [m, m] are written).
extern "C" int sb_pair_intersections(const void* bits, const int* bbox, const int* area, int m, int H, int W,
                                     double area_ratio_thresh, int* inter, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(m > 0 && H > 0 && W > 0, "sb_pair_intersections: empty");
  const long long threads = static_cast<long long>(m) * m * 32;
  long long g = (threads + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  pair_inter_kernel<<<static_cast<int>(g), 256, 0, stream>>>(static_cast<const uint32_t*>(bits), bbox, area, m, H,
                                                             (W + 31) / 32, area_ratio_thresh, inter);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_unpack_bits(const void* bits, const int* sel, int m, int H, int W,
                              unsigned char* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(m > 0 && H > 0 && W > 0, "sb_unpack_bits: empty");
  const long long total = static_cast<long long>(m) * H * W;
  long long g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  unpack_bits_kernel<<<static_cast<int>(g), 256, 0, stream>>>(static_cast<const uint32_t*>(bits), sel, m,
                                                             H, W, (W + 31) / 32, out);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_gather_rows(const void* src, const int* sel, int m, long long row_words, void* dst,
                              void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(m > 0 && row_words > 0, "sb_gather_rows: empty");
  const long long total = static_cast<long long>(m) * row_words;
  long long g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  gather_rows_kernel<<<static_cast<int>(g), 256, 0, stream>>>(static_cast<const uint32_t*>(src), sel, m,
                                                              row_words, static_cast<uint32_t*>(dst));
  SB_CHECK_LAUNCH();
  return SB_OK;
}
