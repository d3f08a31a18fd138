// itd_sift2d.cuh -- SURVEY.md 8f rank 3: data movement of the 2-D "crossways" ensemble ITD.
//
// Reference: siftED2D.ipynb code cell 1 (raw JSON :233-278): crossways_itd_baseline_extract runs the spline
// baseline (itd_baseline_extract, "< 10 extrema -> return x") along every row, along every column, then along the
// rows of the column result and the columns of the row result, and averages the two; the ensemble driver
// retrieve_statistical_image_component feeds it data + v and data - v for 10 noise draws v and averages.
//
// The 1-D passes are the ordinary batched spline level (itd_spline.cuh) on (images * rows) signals; a column pass
// is a row pass on the transposed batch.  The kernels here are the tiled transposes between the passes, with the
// ensemble's element-wise steps fused into them.  HBM-bound copies: 2 s bytes per element each.
#pragma once

#include <cuda_runtime.h>

namespace pyitd {

constexpr int kTrTile = 32;

// out[b, c, r] = in[b, r, c]; grid (ceil(W/32), ceil(H/32), B), block (32, 8)
template <typename T>
__global__ void __launch_bounds__(256) transpose_batch_kernel(const T *__restrict__ in, T *__restrict__ out, int H, int W) {
    __shared__ T tile[kTrTile][kTrTile + 1];
    const long long img = (long long)blockIdx.z * H * W;
    const int c0 = blockIdx.x * kTrTile, r0 = blockIdx.y * kTrTile;
    for (int j = threadIdx.y; j < kTrTile; j += 8) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        if (r < H && c < W) tile[j][threadIdx.x] = in[img + (long long)r * W + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < kTrTile; j += 8) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (r < H && c < W) out[img + (long long)c * H + r] = tile[threadIdx.x][j];
    }
}

// out[b, r, c] = (in_t[b, c, r] + other[b, r, c]) / 2 -- the last column pass comes back transposed and is averaged
// with the row-of-columns result (siftED2D.ipynb cell 1: returna = (lengthwise + crosswise) / 2.0)
template <typename T>
__global__ void __launch_bounds__(256) transpose_average_kernel(const T *__restrict__ in_t, const T *__restrict__ other,
                                                                T *__restrict__ out, int H, int W) {
    __shared__ T tile[kTrTile][kTrTile + 1];
    const long long img = (long long)blockIdx.z * H * W;
    const int c0 = blockIdx.x * kTrTile, r0 = blockIdx.y * kTrTile;
    for (int j = threadIdx.y; j < kTrTile; j += 8) {       // read in_t[b, c0 + j, r0 + x]
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (r < H && c < W) tile[j][threadIdx.x] = in_t[img + (long long)c * H + r];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < kTrTile; j += 8) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        if (r < H && c < W) {
            const long long o = img + (long long)r * W + c;
            out[o] = (T)(((double)tile[threadIdx.x][j] + (double)other[o]) / 2.0);
        }
    }
}

// ensemble members: n[2e] = data + v_e, n[2e + 1] = data - v_e  (v * -1 + data in the reference)
template <typename T>
__global__ void __launch_bounds__(256) ensemble_members_kernel(const T *__restrict__ data, const T *__restrict__ noise,
                                                               T *__restrict__ members, long long hw, int draws) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hw) return;
    const double d = (double)data[i];
    for (int e = 0; e < draws; ++e) {
        const double v = (double)noise[(long long)e * hw + i];
        members[(2ll * e) * hw + i] = (T)(v + d);
        members[(2ll * e + 1) * hw + i] = (T)((v * -1.0) + d);
    }
}

// b_e = (y[2e] + y[2e+1]) / 2 ; lowpass = (b_0 + b_1 + ...) / draws, accumulated in the reference's order
template <typename T>
__global__ void __launch_bounds__(256) ensemble_mean_kernel(const T *__restrict__ y, T *__restrict__ lowpass, long long hw,
                                                            int draws) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hw) return;
    double acc = 0.0;
    for (int e = 0; e < draws; ++e) {
        const double b = ((double)y[(2ll * e) * hw + i] + (double)y[(2ll * e + 1) * hw + i]) / 2.0;
        acc += b;
    }
    lowpass[i] = (T)(acc / ((double)draws * 1.0));
}

}  // namespace pyitd
