// otb_tiles.cu — sparse transport of detector images.
//
// A detector image of an imaging system is a (Ny, Nx, 4) float64 histogram of 29-143 MB of which a few per cent of
// the pixels are non-zero (the double-Gauss workload: 3e4 of 4.5e6).  Moving it whole — the all-reduce over the
// GPUs of a sharded trace, the device -> host copy behind RenderImage.data — costs more than the
// binning that produced it.  These kernels move only the occupied T x T pixel tiles:
//   mask    mask[t] |= tile t holds a non-zero value          (all-reduce MAX of the mask = union over the ranks)
//   pack    header = [count, overflow, tile ids ...]; packed[k] = tile ids[k] copied out (capacity `cap` tiles)
//   unpack  packed tiles written back (after the SUM all-reduce of `packed` over the ranks)
// Every rank packs by the SAME union mask, so slot k means the same tile everywhere and the packed buffers can be
// summed element-wise.  count > cap leaves the image untouched (overflow flag): the caller falls back to the dense
// path.  Edge tiles are padded with zeros.
#include "otb_common.cuh"

__global__ void __launch_bounds__(256) tiles_mask_kernel(const double* __restrict__ img, int Ny, int Nx, int T, int ntx,
                                                         int* __restrict__ mask)
{
    const int t = blockIdx.x, ty = t/ntx, tx = t - ty*ntx;
    bool any = false;
    for (int i = threadIdx.x; i < T*T; i += blockDim.x) {
        const int y = ty*T + i/T, x = tx*T + i%T;
        if (y < Ny && x < Nx) {
            const double4 v0 = *reinterpret_cast<const double4*>(img + 4*((int64_t)y*Nx + x));
            any = any || (v0.x != 0.0) || (v0.y != 0.0) || (v0.z != 0.0) || (v0.w != 0.0);
        }
    }
    if (__syncthreads_or(any) && threadIdx.x == 0) mask[t] = 1;
}

// one block: stable compaction of the tile ids with mask != 0 into header[2..], count into header[0]
__global__ void __launch_bounds__(1024) tiles_list_kernel(const int* __restrict__ mask, int ntiles, int cap, int* __restrict__ header)
{
    __shared__ int warp_sums[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int t0 = 0; t0 < ntiles; t0 += blockDim.x) {
        const int t = t0 + threadIdx.x;
        const int f = (t < ntiles && mask[t] != 0) ? 1 : 0;
        const unsigned b = __ballot_sync(0xffffffffu, f);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) warp_sums[w] = __popc(b);
        __syncthreads();
        int off = base;
        for (int k = 0; k < w; ++k) off += warp_sums[k];
        const int pos = off + __popc(b & ((1u << lane) - 1));
        if (f && pos < cap) header[2 + pos] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            int s = 0;
            for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += warp_sums[k];
            base += s;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        header[0] = base;
        header[1] = base > cap ? 1 : 0;
    }
}

// block k copies tile header[2 + k] (k < count <= cap) between the image and its slot; dir 0: image -> packed,
// 1: packed -> image.  Slots beyond the count are zeroed when packing (they take part in the element-wise reduction).
__global__ void __launch_bounds__(256) tiles_copy_kernel(double* __restrict__ img, int Ny, int Nx, int T, int ntx,
                                                         const int* __restrict__ header, int cap, double* __restrict__ packed, int dir)
{
    const int k = blockIdx.x;
    const int count = header[0];
    double* slot = packed + (int64_t)k*T*T*4;
    if (count > cap) return;                              // overflow: the dense path takes over, nothing is touched
    if (k >= count) {
        if (dir == 0) for (int i = threadIdx.x; i < T*T*4; i += blockDim.x) slot[i] = 0.0;
        return;
    }
    const int t = header[2 + k], ty = t/ntx, tx = t - ty*ntx;
    for (int i = threadIdx.x; i < T*T; i += blockDim.x) {
        const int y = ty*T + i/T, x = tx*T + i%T;
        const bool in = y < Ny && x < Nx;
        double4* s4 = reinterpret_cast<double4*>(slot + 4*(int64_t)i);
        if (dir == 0) *s4 = in ? *reinterpret_cast<const double4*>(img + 4*((int64_t)y*Nx + x)) : make_double4(0.0, 0.0, 0.0, 0.0);
        else if (in) *reinterpret_cast<double4*>(img + 4*((int64_t)y*Nx + x)) = *s4;
    }
}

extern "C" {

int otb_image_tiles_mask(const double* img_d, int32_t Ny, int32_t Nx, int32_t T, int32_t* mask_d, void* stream)
{
    if (!img_d || !mask_d || Ny < 1 || Nx < 1 || T < 1) { otb_set_error("invalid argument"); return OTB_ERR_INVALID_ARG; }
    const int ntx = (Nx + T - 1)/T, nty = (Ny + T - 1)/T;
    tiles_mask_kernel<<<ntx*nty, 256, 0, (cudaStream_t)stream>>>(img_d, Ny, Nx, T, ntx, mask_d);
    OTB_CUDA(cudaGetLastError());
    return OTB_OK;
}

int otb_image_tiles_pack(const double* img_d, int32_t Ny, int32_t Nx, int32_t T, const int32_t* mask_d, int32_t cap,
                         int32_t* header_d, double* packed_d, void* stream)
{
    if (!img_d || !mask_d || !header_d || !packed_d || Ny < 1 || Nx < 1 || T < 1 || cap < 1) { otb_set_error("invalid argument"); return OTB_ERR_INVALID_ARG; }
    const int ntx = (Nx + T - 1)/T, nty = (Ny + T - 1)/T;
    cudaStream_t st = (cudaStream_t)stream;
    tiles_list_kernel<<<1, 1024, 0, st>>>(mask_d, ntx*nty, cap, header_d);
    tiles_copy_kernel<<<cap, 256, 0, st>>>((double*)img_d, Ny, Nx, T, ntx, header_d, cap, packed_d, 0);
    OTB_CUDA(cudaGetLastError());
    return OTB_OK;
}

int otb_image_tiles_unpack(double* img_d, int32_t Ny, int32_t Nx, int32_t T, const int32_t* header_d, int32_t cap,
                           const double* packed_d, void* stream)
{
    if (!img_d || !header_d || !packed_d || Ny < 1 || Nx < 1 || T < 1 || cap < 1) { otb_set_error("invalid argument"); return OTB_ERR_INVALID_ARG; }
    const int ntx = (Nx + T - 1)/T;
    tiles_copy_kernel<<<cap, 256, 0, (cudaStream_t)stream>>>(img_d, Ny, Nx, T, ntx, header_d, cap, (double*)packed_d, 1);
    OTB_CUDA(cudaGetLastError());
    return OTB_OK;
}

}  // extern "C"
