// saber_b200 — 3-D connected components (26-connectivity) with small-component removal and compact
// relabelling, bit-exact with REF saber/segmenters/utils.py:88-131 (`separate_masks`:
// scipy.ndimage.label with a full 3x3x3 structure -> drop components with fewer than min_vol voxels ->
// compact relabel 1..K). scipy numbers components by the raster-scan position of their first voxel; a
// union-find whose root is the component's minimum linear index reproduces exactly that order, so the
// compact relabel is an exclusive prefix count of surviving roots.
//
// Passes (all HBM-streaming, integer): init parent = self | merge with the 13 raster-preceding
// neighbours (atomicMin hooking) | flatten | warp-aggregated component sizes | chunked prefix count of
// surviving roots -> new ids | relabel in place. Algorithmic bytes: 2 B read + 4 B written per voxel.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int CHUNK = 2048;  // voxels per ranking chunk (256 threads x 8)

__device__ __forceinline__ int uf_find(const int* __restrict__ P, int v) {
  int p = P[v];
  while (p != v) {
    v = p;
    p = P[v];
  }
  return v;
}

// find with path halving: every visited node is re-pointed at its grandparent. Parents only ever decrease (atomicMin), so
// a node keeps pointing at a smaller member of its own component whatever the interleaving, and chains cannot cycle.
// Without it the merge pass built chains as long as an object has rows (every row head hooks onto the head above it at the
// same time) and walked them voxel by voxel: 10.7 GB of DRAM reads on a 0.84 GB parent array (profiles/r02zzb).
__device__ __forceinline__ int uf_find_halve(int* P, int v) {
  int p = P[v];
  while (p != v) {
    const int gp = P[p];
    if (gp != p) atomicMin(&P[v], gp);
    v = p;
    p = gp;
  }
  return v;
}

__device__ __forceinline__ void uf_unite(int* P, int a, int b) {
  bool done;
  do {
    a = uf_find_halve(P, a);
    b = uf_find_halve(P, b);
    if (a < b) {
      const int old = atomicMin(&P[b], a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      const int old = atomicMin(&P[a], b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

// parent initialisation: a foreground voxel points at the first voxel of its foreground RUN inside the warp's 32-voxel
// segment of the row (one ballot; runs never cross a row end), background voxels get -1. Chains start one hop from their
// run head, and the merge pass only has to link run heads to the left (at segment boundaries).
template <typename T>
__global__ void __launch_bounds__(256)
ccl_init_kernel(const T* __restrict__ vol, int* __restrict__ P, int* __restrict__ aux, long long n, int X,
                int* __restrict__ chunk_flag) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long n_round = ((n + 31) / 32) * 32;
  const int lane = threadIdx.x & 31;
#pragma unroll 4
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    const bool in = i < n;
    const bool fg = in && vol[i] != 0;
    const bool row_start = in && (i % X) == 0;
    const uint32_t m = __ballot_sync(0xffffffffu, fg);
    const uint32_t brk = __ballot_sync(0xffffffffu, row_start);
    // lanes a run cannot extend across (to lower lanes): background lanes, and the lane just before a row start
    const uint32_t stop = (~m | (brk >> 1)) & ((1u << lane) - 1u);
    const int first = stop ? 32 - __clz(stop) : 0;  // first lane of this lane's run
    if (in) {
      P[i] = fg ? static_cast<int>(i - lane + first) : -1;
      // sizes / ids live at roots only, and a root is always the head of its run: nothing else of aux is ever read
      if (fg && first == lane) aux[i] = 0;
    }
    // 2048-voxel chunks without foreground are skipped by the counting / ranking passes (a warp's 32 voxels lie in one chunk)
    if (m != 0u && lane == 0) chunk_flag[i / CHUNK] = 1;
  }
}

// 26-connectivity merge over the 13 raster-preceding neighbours, with the redundant unions removed:
//  * left neighbour: already linked by the initialisation unless this voxel heads its run (segment boundary);
//  * in each of the 4 preceding rows (previous row of the plane, 3 rows of the previous plane) the cells dx = -1, 0, +1
//    are consecutive voxels of ONE row: if the centre is foreground it is linked to both others by that row's own runs,
//    so one union (with the centre) suffices; otherwise dx = -1 and dx = +1 are united separately;
//  * if the left neighbour v-1 is foreground, everything at dx <= 0 is also a neighbour of v-1 (its dx <= +1) and is
//    linked through it: only dx = +1 remains, and only when the centre is background.
// Inside a solid object a voxel performs no union at all (the voxel-by-voxel form did 13 find/atomicMin pairs: 22 ms of
// the 23 ms a 200 x 1024 x 1024 volume took, profiles/r02zc). Roots are component minima under any union order, so the
// labels are unchanged (scipy's raster numbering).
__global__ void __launch_bounds__(256)
ccl_merge_kernel(int* __restrict__ P, int Z, int Y, int X, long long i0, long long i1) {
  const long long plane = static_cast<long long>(Y) * X;
  for (long long i = i0 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < i1;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (P[i] < 0) continue;
    const unsigned ui = static_cast<unsigned>(i), row_id = ui / static_cast<unsigned>(X);  // n < 2^31: 32-bit divisions
    const int x = static_cast<int>(ui - row_id * static_cast<unsigned>(X));
    const int z = static_cast<int>(row_id / static_cast<unsigned>(Y));
    const int y = static_cast<int>(row_id - static_cast<unsigned>(z) * static_cast<unsigned>(Y));
    const int v = static_cast<int>(i);
    const bool left = x > 0 && P[i - 1] >= 0;
    // a run head with a foreground voxel to its left exists only at a 32-voxel segment boundary (decided from the
    // geometry: P[v] itself may already have been lowered by another thread's union)
    if (left && (i & 31) == 0) uf_unite(P, v, v - 1);
    auto row = [&](long long r) {  // r = linear index of the cell above / behind v (dx = 0) in a preceding row
      if (P[r] >= 0) {  // centre set: the only union that can be needed (none inside a run), and no other cell is read
        if (!left) uf_unite(P, v, static_cast<int>(r));
        return;
      }
      if (!left && x > 0 && P[r - 1] >= 0) uf_unite(P, v, static_cast<int>(r - 1));
      if (x + 1 < X && P[r + 1] >= 0) uf_unite(P, v, static_cast<int>(r + 1));
    };
    if (y > 0) row(i - X);
    if (z > 0) {
      if (y > 0) row(i - plane - X);
      row(i - plane);
      if (y + 1 < Y) row(i - plane + X);
    }
  }
}

// 6-connectivity (scipy.ndimage.label's default structure: face neighbours only) over the 3 raster-preceding neighbours.
// Same run-head initialisation; the up / back union is implied when the left neighbour and ITS up / back neighbour are
// both foreground (they are linked through the rows' own runs).
__global__ void __launch_bounds__(256)
ccl_merge6_kernel(int* __restrict__ P, int Z, int Y, int X, long long i0, long long i1) {
  const long long plane = static_cast<long long>(Y) * X;
  for (long long i = i0 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < i1;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (P[i] < 0) continue;
    const int x = static_cast<int>(i % X);
    const int y = static_cast<int>((i / X) % Y);
    const int z = static_cast<int>(i / plane);
    const int v = static_cast<int>(i);
    const bool left = x > 0 && P[i - 1] >= 0;
    if (left && (i & 31) == 0) uf_unite(P, v, v - 1);
    if (y > 0 && P[i - X] >= 0 && !(left && P[i - X - 1] >= 0)) uf_unite(P, v, static_cast<int>(i - X));
    if (z > 0 && P[i - plane] >= 0 && !(left && P[i - plane - 1] >= 0)) uf_unite(P, v, static_cast<int>(i - plane));
  }
}

// flatten + warp-aggregated size count: aux[root] += #voxels
__global__ void __launch_bounds__(256)
ccl_flatten_count_kernel(int* __restrict__ P, int* __restrict__ aux, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long n_round = ((n + 31) / 32) * 32;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    int root = -1;
    if (i < n && P[i] >= 0) {
      root = uf_find(P, static_cast<int>(i));
      P[i] = root;
    }
    const uint32_t active = __ballot_sync(0xffffffffu, root >= 0);
    if (root >= 0) {
      const uint32_t peers = __match_any_sync(active, root);
      if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&aux[root], __popc(peers));
    }
  }
}

// per-chunk count of surviving roots (P[v] == v and size >= min_vol)
__global__ void __launch_bounds__(256)
ccl_chunk_count_kernel(const int* __restrict__ P, const int* __restrict__ aux, long long n, int min_vol,
                       int* __restrict__ chunk_cnt) {
  if (chunk_cnt[blockIdx.x] == 0) return;  // no foreground in this chunk (flag written by the init pass): count stays 0
  const long long base = static_cast<long long>(blockIdx.x) * CHUNK;
  int c = 0;
  if (base + CHUNK <= n && (reinterpret_cast<uintptr_t>(P) & 15) == 0) {  // whole chunk: 8 parents per thread, two 128-bit loads
    const int4* p4 = reinterpret_cast<const int4*>(P + base);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = (h * 256 + threadIdx.x) * 4;
      const int4 q = p4[h * 256 + threadIdx.x];
      const int v0 = static_cast<int>(base) + k;
      if (q.x == v0 && aux[v0] >= min_vol) ++c;
      if (q.y == v0 + 1 && aux[v0 + 1] >= min_vol) ++c;
      if (q.z == v0 + 2 && aux[v0 + 2] >= min_vol) ++c;
      if (q.w == v0 + 3 && aux[v0 + 3] >= min_vol) ++c;
    }
  } else {
    for (int k = threadIdx.x; k < CHUNK; k += 256) {
      const long long v = base + k;
      if (v < n && P[v] == static_cast<int>(v) && aux[v] >= min_vol) ++c;
    }
  }
  __shared__ int s[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < 8; ++k) t += s[k];
    chunk_cnt[blockIdx.x] = t;
  }
}

// in-place exclusive scan of chunk_cnt[0..nchunks) by a single block; total -> *total_out
__global__ void __launch_bounds__(1024)
ccl_scan_kernel(int* __restrict__ chunk_cnt, int nchunks, int* __restrict__ total_out) {
  __shared__ int warp_tot[32];
  __shared__ int running;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i0 = 0; i0 < nchunks; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    const int v = i < nchunks ? chunk_cnt[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int off = running;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (i < nchunks) chunk_cnt[i] = off + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) running = off + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = running;
}

// aux[root] <- -(compact id) (1-based) for surviving roots; dropped roots keep their (positive) size, so only chunks that
// hold a surviving root are touched at all (chunk_off[c + 1] == chunk_off[c] otherwise; chunk_off[nchunks] = total)
__global__ void __launch_bounds__(256)
ccl_assign_kernel(const int* __restrict__ P, int* __restrict__ aux, long long n, int min_vol,
                  const int* __restrict__ chunk_off, int* __restrict__ sizes_out) {
  if (chunk_off[blockIdx.x + 1] == chunk_off[blockIdx.x]) return;
  const long long base = static_cast<long long>(blockIdx.x) * CHUNK;
  __shared__ int warp_tot[8];
  __shared__ int running;
  if (threadIdx.x == 0) running = chunk_off[blockIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k0 = 0; k0 < CHUNK; k0 += 256) {
    const long long v = base + k0 + threadIdx.x;
    const bool is_root = v < n && P[v] == static_cast<int>(v);
    const bool keep = is_root && aux[v] >= min_vol;
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = running;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (keep) {
      const int id = off + __popc(bal & ((1u << lane) - 1u)) + 1;
      if (sizes_out) sizes_out[id - 1] = aux[v];  // voxel count of component `id`
      aux[v] = -id;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 8; ++w) t += warp_tot[w];
      running += t;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
ccl_relabel_kernel(int* __restrict__ P, const int* __restrict__ aux, long long n) {
#pragma unroll 4
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = P[i];
    int lab = 0;
    if (r >= 0) {
      const int a = aux[r];  // -(id) for a surviving root, the (positive) size of a dropped one
      if (a < 0) lab = -a;
    }
    P[i] = lab;
  }
}

}  // namespace

// vol: [Z, Y, X] (elem_bytes 1, 2 or 4; non-zero = foreground). labels: [Z, Y, X] uint32 output (also the
// union-find parent array). aux: workspace of Z*Y*X int32. chunk_ws: workspace of ceil(Z*Y*X / 2048) + 1
// int32; its last element receives the number of components kept. Components with fewer than min_vol
// voxels are removed (min_vol <= 1 keeps everything).
static int ccl3d_run(const void* vol, int elem_bytes, int Z, int Y, int X, int min_vol, int conn, void* labels, int* aux,
                     int* chunk_ws, int* sizes_out, void* stream_);

extern "C" int sb_ccl3d_26(const void* vol, int elem_bytes, int Z, int Y, int X, int min_vol, void* labels,
                           int* aux, int* chunk_ws, void* stream_) {
  return ccl3d_run(vol, elem_bytes, Z, Y, X, min_vol, 26, labels, aux, chunk_ws, nullptr, stream_);
}

// conn: 6 (scipy.ndimage.label default) or 26. sizes_out (optional, one int per surviving component, at least as many
// entries as components can exist: Z*Y*X / min_vol + 1 is always enough) receives the voxel count of component id at
// sizes_out[id - 1]. REF saber/analysis/refine_membranes.py:136-249 (_remove_small_objects / _get_largest_component).
extern "C" int sb_ccl3d(const void* vol, int elem_bytes, int Z, int Y, int X, int min_vol, int conn, void* labels,
                        int* aux, int* chunk_ws, int* sizes_out, void* stream_) {
  SB_REQUIRE(conn == 6 || conn == 26, "sb_ccl3d: connectivity must be 6 or 26 (got %d)", conn);
  return ccl3d_run(vol, elem_bytes, Z, Y, X, min_vol, conn, labels, aux, chunk_ws, sizes_out, stream_);
}

static int ccl3d_run(const void* vol, int elem_bytes, int Z, int Y, int X, int min_vol, int conn, void* labels, int* aux,
                     int* chunk_ws, int* sizes_out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(Z > 0 && Y > 0 && X > 0, "sb_ccl3d_26: empty volume");
  const long long n = static_cast<long long>(Z) * Y * X;
  SB_REQUIRE(n < (1ll << 31), "sb_ccl3d_26: volume too large for int32 indices (%lld voxels)", n);
  SB_REQUIRE(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4, "sb_ccl3d_26: elem_bytes must be 1, 2 or 4");
  int* P = static_cast<int*>(labels);
  long long g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  const int grid = static_cast<int>(g);
  const int nchunks_ = static_cast<int>((n + CHUNK - 1) / CHUNK);
  SB_CHECK_CUDA(cudaMemsetAsync(chunk_ws, 0, sizeof(int) * (static_cast<size_t>(nchunks_) + 1), stream));
  if (elem_bytes == 1)
    ccl_init_kernel<unsigned char><<<grid, 256, 0, stream>>>(static_cast<const unsigned char*>(vol), P, aux, n, X, chunk_ws);
  else if (elem_bytes == 2)
    ccl_init_kernel<unsigned short><<<grid, 256, 0, stream>>>(static_cast<const unsigned short*>(vol), P, aux, n, X, chunk_ws);
  else
    ccl_init_kernel<unsigned int><<<grid, 256, 0, stream>>>(static_cast<const unsigned int*>(vol), P, aux, n, X, chunk_ws);
  SB_CHECK_LAUNCH();
  // The merge pass runs slab by slab (about 32 MB of parents per launch): within one launch the CTAs drift apart (a CTA
  // full of foreground is much slower than one of background), and over the whole volume that drift made every neighbour
  // read a DRAM miss (10 GB read for a 0.84 GB parent array, profiles/r02zzb). A slab and the plane before it stay in L2.
  {
    static const long long slab_mi = [] { const char* e = getenv("SB_CCL_SLAB"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 8; }();
    const long long slab = slab_mi << 20;  // SB_CCL_SLAB: Mi voxels per merge launch (measurement switch)
    for (long long i0 = 0; i0 < n; i0 += slab) {
      const long long i1 = i0 + slab < n ? i0 + slab : n;
      long long gl = (i1 - i0 + 255) / 256;
      if (gl > 148 * 16) gl = 148 * 16;
      const int g = static_cast<int>(gl);
      if (conn == 26)
        ccl_merge_kernel<<<g, 256, 0, stream>>>(P, Z, Y, X, i0, i1);
      else
        ccl_merge6_kernel<<<g, 256, 0, stream>>>(P, Z, Y, X, i0, i1);
      SB_CHECK_LAUNCH();
    }
  }
  ccl_flatten_count_kernel<<<grid, 256, 0, stream>>>(P, aux, n);
  SB_CHECK_LAUNCH();
  const int nchunks = static_cast<int>((n + CHUNK - 1) / CHUNK);
  const int mv = min_vol > 1 ? min_vol : 1;
  ccl_chunk_count_kernel<<<nchunks, 256, 0, stream>>>(P, aux, n, mv, chunk_ws);
  SB_CHECK_LAUNCH();
  ccl_scan_kernel<<<1, 1024, 0, stream>>>(chunk_ws, nchunks, chunk_ws + nchunks);
  SB_CHECK_LAUNCH();
  ccl_assign_kernel<<<nchunks, 256, 0, stream>>>(P, aux, n, mv, chunk_ws, sizes_out);
  SB_CHECK_LAUNCH();
  ccl_relabel_kernel<<<grid, 256, 0, stream>>>(P, aux, n);
  SB_CHECK_LAUNCH();
  return SB_OK;
}
