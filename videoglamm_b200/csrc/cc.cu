// 8-connected component labelling + per-pixel component area for binary masks (sm_100a). The B200-side equivalent of the reference's only native
// extension, sam2._C.get_connected_componnets (sam2/csrc/connected_components.cu:213-289): same inputs, same two outputs, same label VALUES.
//
// Output contract of the reference (derived from its kernels :59-201): label of a foreground pixel = 1 + min over its component of the linear
// index (row * W + col) of the TOP-LEFT pixel of the 2x2 block (even row, even col) each member pixel lies in; background = 0. counts = number
// of foreground pixels of the component at every foreground pixel, 0 on the background. H and W must be even.
//
// Design (not the reference's block-based union-find): one lock-free union-find over PIXELS that are renumbered block-major,
//   id(row, col) = ((row >> 1) * (W >> 1) + (col >> 1)) * 4 + (row & 1) * 2 + (col & 1),
// so that the minimum id of a component lies in its minimum 2x2 block and the root of the union-find (roots are minima) directly encodes the
// reference's label. All N images are processed by the same 4 launches (the reference issues 6 launches PER image):
//   init (parent = id on foreground, -1 on background; area = 0) -> merge (union with the W, NW, N, NE neighbours) ->
//   count (find root with path compression, atomicAdd area[root]) -> final (labels / counts in pixel order).
// HBM-bound integer work: ~ (1 + 4 + 4) B read + 8 B written per pixel over the 4 passes, plus the 8 B/pixel workspace.
#include "common.h"
#include "pdl.cuh"

namespace vg {

__device__ __forceinline__ int cc_id(int row, int col, int w) { return ((row >> 1) * (w >> 1) + (col >> 1)) * 4 + (row & 1) * 2 + (col & 1); }

__device__ __forceinline__ int cc_find(const int* parent, int a) {
  int p = parent[a];
  while (p != a) {
    a = p;
    p = parent[a];
  }
  return a;
}

__device__ __forceinline__ void cc_union(int* parent, int a, int b) {
  while (true) {
    a = cc_find(parent, a);
    b = cc_find(parent, b);
    if (a == b) return;
    if (a > b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(parent + b, a);  // hook the larger root under the smaller one
    if (old == b) return;
    b = old;  // somebody else re-parented b meanwhile: retry from there
  }
}

__global__ void cc_init_kernel(const uint8_t* __restrict__ mask, int* __restrict__ parent, int* __restrict__ area, int n, int h, int w) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long hw = static_cast<long long>(h) * w;
  if (i >= n * hw) return;
  const int img = static_cast<int>(i / hw), p = static_cast<int>(i - img * hw);
  const int row = p / w, col = p - row * w;
  const int id = cc_id(row, col, w);
  parent[img * hw + id] = mask[i] ? id : -1;
  area[img * hw + id] = 0;
}

__global__ void cc_merge_kernel(const uint8_t* __restrict__ mask, int* __restrict__ parent, int n, int h, int w) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long hw = static_cast<long long>(h) * w;
  if (i >= n * hw) return;
  if (!mask[i]) return;
  const int img = static_cast<int>(i / hw), p = static_cast<int>(i - img * hw);
  const int row = p / w, col = p - row * w;
  const uint8_t* m = mask + img * hw;
  int* par = parent + img * hw;
  const int id = cc_id(row, col, w);
  if (col > 0 && m[p - 1]) cc_union(par, id, cc_id(row, col - 1, w));
  if (row > 0) {
    if (col > 0 && m[p - w - 1]) cc_union(par, id, cc_id(row - 1, col - 1, w));
    if (m[p - w]) cc_union(par, id, cc_id(row - 1, col, w));
    if (col + 1 < w && m[p - w + 1]) cc_union(par, id, cc_id(row - 1, col + 1, w));
  }
}

__global__ void cc_count_kernel(int* __restrict__ parent, int* __restrict__ area, int n, long long hw) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n * hw) return;
  const long long base = (i / hw) * hw;
  const int id = static_cast<int>(i - base);
  if (parent[i] < 0) return;
  const int root = cc_find(parent + base, id);
  parent[i] = root;  // roots are final after the merge pass: a plain store is a valid path compression
  atomicAdd(area + base + root, 1);
}

__global__ void cc_final_kernel(const uint8_t* __restrict__ mask, const int* __restrict__ parent, const int* __restrict__ area, int* __restrict__ labels,
                                int* __restrict__ counts, int n, int h, int w) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long hw = static_cast<long long>(h) * w;
  if (i >= n * hw) return;
  int lab = 0, cnt = 0;
  if (mask[i]) {
    const int img = static_cast<int>(i / hw), p = static_cast<int>(i - img * hw);
    const int row = p / w, col = p - row * w;
    const int root = parent[img * hw + cc_id(row, col, w)];
    const int blk = root >> 2, half_w = w >> 1;
    lab = (blk / half_w) * 2 * w + (blk % half_w) * 2 + 1;
    cnt = area[img * hw + root];
  }
  labels[i] = lab;
  counts[i] = cnt;
}

}  // namespace vg

using namespace vg;

extern "C" int vg_connected_components(const unsigned char* mask, int* labels, int* counts, int n, int h, int w, void* workspace,
                                       long long workspace_bytes, void* stream) {
  if (!mask || !labels || !counts || n <= 0 || h <= 0 || w <= 0) return set_error(VG_EINVAL, "vg_connected_components: bad argument");
  if ((h % 2) || (w % 2)) return set_error(VG_EINVAL, "vg_connected_components: height and width must be even (got %d x %d)", h, w);
  const long long total = static_cast<long long>(n) * h * w;
  if (static_cast<long long>(h) * w >= (1LL << 31) - 1) return set_error(VG_ENOSUP, "vg_connected_components: image too large");
  if (!workspace || workspace_bytes < 2 * total * 4) return set_error(VG_EINVAL, "vg_connected_components: workspace of %lld bytes needed", 2 * total * 4);
  int* parent = static_cast<int*>(workspace);
  int* area = parent + total;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  launch_k(cc_init_kernel, dim3(blocks), dim3(256), 0, s, mask, parent, area, n, h, w);
  launch_k(cc_merge_kernel, dim3(blocks), dim3(256), 0, s, mask, parent, n, h, w);
  launch_k(cc_count_kernel, dim3(blocks), dim3(256), 0, s, parent, area, n, static_cast<long long>(h) * w);
  launch_k(cc_final_kernel, dim3(blocks), dim3(256), 0, s, mask, parent, area, labels, counts, n, h, w);
  for (int i = 0; i < 4; ++i) count_launch();
  return check_launch("vg_connected_components");
}
