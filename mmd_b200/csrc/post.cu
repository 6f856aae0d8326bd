// Post-sampling integer / host-replacement kernels ("next" rows of SURVEY 8f): batched conflict detection with
// densification and Savitzky-Golay smoothing.  Compiled with -fmad=false like guide.cu: the conflict test must round
// exactly like the reference's op-by-op fp32 expressions (bit-exact integer output).
//
// Reference: mmd/planners/multi_agent/cbs.py:166-246 (CBS.get_conflicts), :446-458 (one get_conflicts call per free sample
// in the 'least_collisions' strategy), mmd/common/trajectory_utils.py:31-38 (smooth_trajs, scipy savgol_filter), :54-69
// (densify_trajs), TR/robots/robot_planar_disk.py:173-203 (check_rr_collisions).
#include "common.cuh"

namespace mmdk {

// densify_trajs (trajectory_utils.py:54-69): point j of segment i = x_i + (j * (x_{i+1} - x_i)) / n, j = 0 .. n-1, plus the
// last waypoint; evaluated on the fly with the reference's operation order.
__device__ __forceinline__ float2 dense_point(const float* __restrict__ path, int T, int n, int td) {
  const int i = td / n, j = td - i * n;
  const float x0 = path[2 * i], y0 = path[2 * i + 1];
  if (j == 0) return make_float2(x0, y0);
  const float x1 = path[2 * i + 2], y1 = path[2 * i + 3];
  const float fj = (float)j, fn = (float)n;
  return make_float2(x0 + (fj * (x1 - x0)) / fn, y0 + (fj * (y1 - y0)) / fn);
}

// One candidate joint state per blockIdx.y: robot `agent` follows cand[c], every other robot r follows base[r].
// coll [n_cand, Td, R, R] uint8 (1 = ||p_a - p_b|| < margin, a != b), count[c] = number of set entries = len(conflict list
// of PointConflicts), dense [n_cand, R, Td, 2] optional (the densified positions the conflicts refer to).
__global__ void conflicts_kernel(const float* __restrict__ base, const float* __restrict__ cand, int agent, int R, int T, int n,
                                 int Td, float margin, uint8_t* __restrict__ coll, int* __restrict__ count,
                                 float* __restrict__ dense) {
  extern __shared__ float2 s_pos[];   // [R] positions of every robot at this dense time step
  const int c = blockIdx.y;
  int local = 0;
  for (int td = blockIdx.x; td < Td; td += gridDim.x) {
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
      const float* path = (r == agent && cand) ? cand + (size_t)c * T * 2 : base + (size_t)r * T * 2;
      const float2 p = dense_point(path, T, n, td);
      s_pos[r] = p;
      if (dense) {
        dense[(((size_t)c * R + r) * Td + td) * 2] = p.x;
        dense[(((size_t)c * R + r) * Td + td) * 2 + 1] = p.y;
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < R * R; e += blockDim.x) {
      const int a = e / R, b = e - a * R;
      const float2 pa = s_pos[a], pb = s_pos[b];
      const float dx = pa.x - pb.x, dy = pa.y - pb.y;
      const float nrm = sqrtf(dx * dx + dy * dy);
      const int hit = (nrm < margin) && (a != b);
      coll[(((size_t)c * Td + td) * R + a) * R + b] = (uint8_t)hit;
      local += hit;
    }
    __syncthreads();
  }
  if (count) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count + c, local);
  }
}

// y[b, h, d] = sum_j S[h, j] x[b, j, d]: the Savitzky-Golay filter INCLUDING scipy's polynomial edge handling is one
// linear operator S [H, H] (built by the host in float64, mmd_b200/smoothing.py); accumulation in double, result fp32.
__global__ void smooth_kernel(const double* __restrict__ S, const float* __restrict__ x, int B, int H, int D,
                              float* __restrict__ y) {
  extern __shared__ double s_S[];   // [H][H] when it fits (H <= 72); longer horizons (multi-tile paths) read S through L2
  const bool staged = H <= 72;
  if (staged) {
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) s_S[i] = S[i];
    __syncthreads();
  }
  const int per = H * D;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < (int64_t)B * per; idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / per), r = (int)(idx - (int64_t)b * per);
    const int h = r / D, d = r - h * D;
    const float* xb = x + (size_t)b * per + d;
    const double* row = (staged ? s_S : S) + (size_t)h * H;
    double acc = 0.0;
    for (int j = 0; j < H; ++j) acc += row[j] * (double)xb[(size_t)j * D];
    y[idx] = (float)acc;
  }
}

}  // namespace mmdk

using namespace mmdk;

extern "C" {

int mmdk_get_conflicts(const float* base_paths_dev, const float* cand_paths_dev, int n_cand, int agent_id, int R, int T,
                       int densify, float margin, uint8_t* coll_dev, int32_t* count_dev, float* dense_dev, void* stream) {
  if (!base_paths_dev || !coll_dev) return fail(MMDK_EINVAL, "null argument");
  if (R < 1 || T < 1 || densify < 1 || n_cand < 1) return fail(MMDK_EINVAL, "R, T, densify and n_cand must be >= 1");
  if (cand_paths_dev && (agent_id < 0 || agent_id >= R)) return fail(MMDK_EINVAL, "agent_id out of range");
  if (!cand_paths_dev && n_cand != 1) return fail(MMDK_EINVAL, "n_cand must be 1 without candidate paths");
  const int Td = (T - 1) * densify + 1;
  if ((size_t)R * sizeof(float2) > 48 * 1024) return fail(MMDK_EINVAL, "too many robots");
  cudaStream_t s = (cudaStream_t)stream;
  if (count_dev) MMDK_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(int32_t) * n_cand, s));
  dim3 grid((unsigned)(Td < 64 ? Td : 64), (unsigned)n_cand);
  conflicts_kernel<<<grid, 256, sizeof(float2) * R, s>>>(base_paths_dev, cand_paths_dev, agent_id, R, T, densify, Td, margin,
                                                        coll_dev, count_dev, dense_dev);
  return check_cuda(cudaGetLastError(), "conflicts_kernel");
}

int mmdk_smooth_trajs(const double* filter_dev, const float* trajs_dev, int B, int H, int D, float* out_dev, void* stream) {
  if (!filter_dev || !trajs_dev || !out_dev) return fail(MMDK_EINVAL, "null argument");
  if (B <= 0) return MMDK_OK;
  if (H < 1 || H > 4096 || D < 1) return fail(MMDK_EINVAL, "horizon must be in [1, 4096]");
  const int64_t n = (int64_t)B * H * D;
  const int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  smooth_kernel<<<grid, 256, H <= 72 ? sizeof(double) * H * H : 0, (cudaStream_t)stream>>>(filter_dev, trajs_dev, B, H, D, out_dev);
  return check_cuda(cudaGetLastError(), "smooth_kernel");
}

}  // extern "C"
