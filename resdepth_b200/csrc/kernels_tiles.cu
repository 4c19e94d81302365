// On-device training-tile producer: DsmOrthoDataset.__getitem__ of the reference (lib/DsmOrthoDataset.py:161-291,
// training strategy) for a whole batch, on rasters resident in HBM -- crop at (y, x), per-tile masked mean-centring
// and division by sigma (lib/DsmOrthoDataset.py:191-210, lib/data_normalization.py:6-26), ortho-image gather and
// normalisation (:213-255), loss mask (:433-470), rot90 / flipud / fliplr augmentation
// (lib/torch_transforms.py:15-157).  The random decisions (positions, image pair, permutation, k, flips) are
// inputs: the host draws them exactly as the reference's sampler would.
#include "common.cuh"

namespace rd {

// per tile: mean of the DSM patch over pixels != nodata, mean of the selected ortho patches
__global__ void __launch_bounds__(256)
tile_means_kernel(const float* __restrict__ dsm_in, const float* __restrict__ orthos, int cols, size_t plane_sz,
                  const int32_t* __restrict__ pos, const int32_t* __restrict__ views, int T, int n_ortho, float nodata,
                  float dsm_mean_in, float ortho_mean_in, float* __restrict__ means /*[n][2]*/) {
  __shared__ double r1[256], r2[256], r3[256];
  const int t = blockIdx.x;
  const int y = pos[2 * t], x = pos[2 * t + 1];
  double ds = 0.0, dc = 0.0, os = 0.0;
  for (int i = threadIdx.x; i < T * T; i += 256) {
    const int r = i / T, c = i - r * T;
    const size_t o = (size_t)(y + r) * cols + x + c;
    const float v = dsm_in[o];
    if (v != nodata) { ds += (double)v; dc += 1.0; }
    for (int k = 0; k < n_ortho; ++k) os += (double)orthos[(size_t)views[t * n_ortho + k] * plane_sz + o];
  }
  r1[threadIdx.x] = ds; r2[threadIdx.x] = dc; r3[threadIdx.x] = os;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      r1[threadIdx.x] += r1[threadIdx.x + s]; r2[threadIdx.x] += r2[threadIdx.x + s]; r3[threadIdx.x] += r3[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    means[2 * t] = isnan(dsm_mean_in) ? (float)(r1[0] / r2[0]) : dsm_mean_in;
    means[2 * t + 1] = isnan(ortho_mean_in) ? (n_ortho ? (float)(r3[0] / ((double)T * T * n_ortho)) : 0.f) : ortho_mean_in;
  }
}

// output plane p of tile t: 0 = loss mask, 1 = target, 2.. = network input channels
__global__ void __launch_bounds__(256)
tile_gather_kernel(const float* __restrict__ dsm_in, const float* __restrict__ dsm_gt, const float* __restrict__ orthos,
                   int cols, size_t plane_sz, const int32_t* __restrict__ pos, const int32_t* __restrict__ views,
                   const int32_t* __restrict__ aug, int T, int n_ortho, int include_dsm, float nodata, float dsm_std,
                   float ortho_std, const float* __restrict__ means, float* __restrict__ input, float* __restrict__ target,
                   uint8_t* __restrict__ mask, float* __restrict__ dsm_mean_out) {
  const int t = blockIdx.z, plane = blockIdx.y;
  const int C = n_ortho + (include_dsm ? 1 : 0);
  const int y = pos[2 * t], x = pos[2 * t + 1];
  const int k = aug[3 * t], vflip = aug[3 * t + 1], hflip = aug[3 * t + 2];
  const float dmean = means[2 * t], omean = means[2 * t + 1];
  if (plane == 0 && blockIdx.x == 0 && threadIdx.x == 0) dsm_mean_out[t] = dmean;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < T * T; i += gridDim.x * 256) {
    const int oi = i / T, oj = i - oi * T;
    // invert  out = fliplr?(flipud?(rot90^k(src)))
    const int q = hflip ? T - 1 - oj : oj;
    const int p = vflip ? T - 1 - oi : oi;
    int si, sj;
    switch (k & 3) {
      case 0: si = p; sj = q; break;
      case 1: si = q; sj = T - 1 - p; break;
      case 2: si = T - 1 - p; sj = T - 1 - q; break;
      default: si = T - 1 - q; sj = p; break;
    }
    const size_t o = (size_t)(y + si) * cols + x + sj;
    const size_t oo = (size_t)t * T * T + i;
    if (plane == 0) {
      const float g = dsm_gt[o];
      mask[oo] = (g != 0.f && g != nodata) ? 1 : 0;
    } else if (plane == 1) {
      target[oo] = __fdiv_rn(__fsub_rn(dsm_gt[o], dmean), dsm_std);
    } else {
      const int c = plane - 2;
      float v;
      if (include_dsm && c == 0) v = __fdiv_rn(__fsub_rn(dsm_in[o], dmean), dsm_std);
      else v = __fdiv_rn(__fsub_rn(orthos[(size_t)views[t * n_ortho + (c - (include_dsm ? 1 : 0))] * plane_sz + o], omean), ortho_std);
      input[((size_t)t * C + c) * T * T + i] = v;
    }
  }
}

int launch_make_tiles(const float* dsm_in, const float* dsm_gt, const float* orthos, int rows, int cols, int nvt,
                      const int32_t* pos, const int32_t* views, const int32_t* aug, int n, int T, int n_ortho,
                      int include_dsm, float nodata, float dsm_std, float ortho_std, float dsm_mean_in,
                      float ortho_mean_in, float* input, float* target, uint8_t* mask, float* dsm_mean_out,
                      float* scratch, cudaStream_t s) {
  if (n <= 0) return 0;
  if (T < 1 || T > rows || T > cols) return fail("make_tiles: tile %d does not fit the %dx%d raster", T, rows, cols);
  (void)nvt;
  if (n_ortho > 0 && (!orthos || !views)) return fail("make_tiles: ortho images requested but not provided");
  if (!include_dsm && n_ortho == 0) return fail("make_tiles: no input channels selected");
  tile_means_kernel<<<n, 256, 0, s>>>(dsm_in, orthos, cols, (size_t)rows * cols, pos, views, T, n_ortho, nodata, dsm_mean_in,
                                      ortho_mean_in, scratch);
  RD_LAUNCHED();
  const int C = n_ortho + (include_dsm ? 1 : 0);
  dim3 grid(cdiv((long long)T * T, 256 * 4), C + 2, n);
  tile_gather_kernel<<<grid, 256, 0, s>>>(dsm_in, dsm_gt, orthos, cols, (size_t)rows * cols, pos, views, aug, T, n_ortho, include_dsm,
                                          nodata, dsm_std, ortho_std, scratch, input, target, mask, dsm_mean_out);
  RD_LAUNCHED();
  return 0;
}

}  // namespace rd
