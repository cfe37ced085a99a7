// Ray generation on the device: the NumPy pinhole generator of dataset.py:109-145, the LLFF
// NDC variant of dataset.py:364-387 and intern/ray.py:59-79 (convert_to_ndc), one thread per pixel.  Neighbour
// rays needed for the radii are recomputed in registers instead of being read back, so the kernel only writes
// (48 B/ray) and the host never materialises or uploads rays.  Compiled with --fmad=false: every fp32 operation
// rounds like the reference's unfused NumPy ops; the radii scale 2/sqrt(12) is applied in fp64 as NumPy does.
#include "common.cuh"

namespace mip360 {

struct RayGenParams {
  const float* c2w;  // [n_img, c2w_rows, 4] row-major, c2w_rows >= 3
  int c2w_rows, n_img, H, W;
  float focal, near, far;
  int ndc;
  float ndc_near;
  long long ray_begin, ray_count;  // the slab [ray_begin, ray_begin + ray_count) of the flattened ray index
};

__device__ __forceinline__ void pinhole_dir(const float* R, int x, int y, int W, int H, float focal, float d[3]) {
  // dataset.py:113-123
  const float cx = ((float)x - (float)W * 0.5f + 0.5f) / focal;
  const float cy = -(((float)y - (float)H * 0.5f + 0.5f)) / focal;
  const float cz = -1.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = (cx * R[i * 4 + 0] + cy * R[i * 4 + 1]) + cz * R[i * 4 + 2];
}

// intern/ray.py:59-79 on one ray
__device__ __forceinline__ void to_ndc(const float o_in[3], const float d[3], float focal, int W, int H, float near,
                                       float o[3], float dn[3]) {
  const float t = -(near + o_in[2]) / (d[2] + 1e-15f);
  float os[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) os[i] = o_in[i] + t * d[i];
  const float fw = -((2.f * focal) / (float)W), fh = -((2.f * focal) / (float)H);
  const float oz = os[2] + 1e-15f;
  o[0] = fw * (os[0] / oz);
  o[1] = fh * (os[1] / oz);
  o[2] = 1.f + 2.f * near / oz;
  const float dz = d[2] + 1e-15f;
  dn[0] = fw * (d[0] / dz - os[0] / oz);
  dn[1] = fh * (d[1] / dz - os[1] / oz);
  dn[2] = -2.f * near / oz;
}

__device__ __forceinline__ float dist3(const float a[3], const float b[3]) {
  const float e0 = a[0] - b[0], e1 = a[1] - b[1], e2 = a[2] - b[2];
  return sqrtf((e0 * e0 + e1 * e1) + e2 * e2);
}

__global__ void __launch_bounds__(256)
raygen_kernel(const RayGenParams p, float* __restrict__ origins, float* __restrict__ directions,
              float* __restrict__ viewdirs, float* __restrict__ radii, float* __restrict__ near_out,
              float* __restrict__ far_out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // output slot
  if (e >= p.ray_count) return;
  const long long ray = p.ray_begin + e;
  const int x = (int)(ray % p.W);
  const long long r = ray / p.W;
  const int y = (int)(r % p.H), img = (int)(r / p.H);
  const float* R = p.c2w + (long long)img * p.c2w_rows * 4;
  const float cam_o[3] = {R[3], R[7], R[11]};

  float d[3];
  pinhole_dir(R, x, y, p.W, p.H, p.focal, d);
  const float nrm = sqrtf((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
  float o_out[3], d_out[3], rad;
  // neighbour pair of dataset.py:128-129 / 370-374: rows (yy, yy+1).  The reference pads the last row with
  // dx[:, -2:-1], i.e. with the pair (H-3, H-2) - not the nearest one - and likewise for the last column
  const int yy = (y < p.H - 1) ? y : p.H - 3, xx = (x < p.W - 1) ? x : p.W - 3;
  if (!p.ndc) {
    float da[3], db[3];
    pinhole_dir(R, x, yy, p.W, p.H, p.focal, da);
    pinhole_dir(R, x, yy + 1, p.W, p.H, p.focal, db);
    rad = (float)((double)dist3(da, db) * 2.0 / sqrt(12.0));
#pragma unroll
    for (int i = 0; i < 3; ++i) { o_out[i] = cam_o[i]; d_out[i] = d[i]; }
  } else {
    to_ndc(cam_o, d, p.focal, p.W, p.H, p.ndc_near, o_out, d_out);
    // dataset.py:369-377: NDC-origin distance to the row neighbour (dx) and to the column neighbour (dy)
    float da[3], oa[3], ob[3], tmp[3];
    pinhole_dir(R, x, yy, p.W, p.H, p.focal, da);
    to_ndc(cam_o, da, p.focal, p.W, p.H, p.ndc_near, oa, tmp);
    pinhole_dir(R, x, yy + 1, p.W, p.H, p.focal, da);
    to_ndc(cam_o, da, p.focal, p.W, p.H, p.ndc_near, ob, tmp);
    const float dx = dist3(oa, ob);
    pinhole_dir(R, xx, y, p.W, p.H, p.focal, da);
    to_ndc(cam_o, da, p.focal, p.W, p.H, p.ndc_near, oa, tmp);
    pinhole_dir(R, xx + 1, y, p.W, p.H, p.focal, da);
    to_ndc(cam_o, da, p.focal, p.W, p.H, p.ndc_near, ob, tmp);
    const float dy = dist3(oa, ob);
    rad = (float)((double)(0.5f * (dx + dy)) * 2.0 / sqrt(12.0));
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    origins[e * 3 + i] = o_out[i];
    directions[e * 3 + i] = d_out[i];
    viewdirs[e * 3 + i] = d[i] / nrm;  // view directions stay the pinhole ones in both variants (dataset.py:383)
  }
  radii[e] = rad;
  near_out[e] = p.near;
  far_out[e] = p.far;
}

// intern/utils.py:17-21 (to8b): (255 * clip(nan_to_num(x), 0, 1)).astype(uint8) — truncation, as NumPy's cast
__device__ __forceinline__ uint32_t to8b_one(float x) {
  return (uint32_t)(uint8_t)(255.f * fminf(fmaxf(nan_to_num_f(x), 0.f), 1.f));
}
// 16 values per thread: four 16-byte loads, one 16-byte store (scalar tail / unaligned buffers: one value per thread)
__global__ void __launch_bounds__(256) to8b_kernel(const float* __restrict__ x, long long n, uint8_t* __restrict__ out,
                                                   int vec) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    const long long i0 = t * 16;
    if (i0 + 16 <= n) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(x + i0 + 4 * j);
        w[j] = to8b_one(v.x) | (to8b_one(v.y) << 8) | (to8b_one(v.z) << 16) | (to8b_one(v.w) << 24);
      }
      *reinterpret_cast<uint4*>(out + i0) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      for (long long i = i0; i < n; ++i) out[i] = (uint8_t)to8b_one(x[i]);
    }
  } else if (t < n) {
    out[t] = (uint8_t)to8b_one(x[t]);
  }
}

}  // namespace mip360

using namespace mip360;

extern "C" int mip360_to8b(const float* x, long long n, uint8_t* out, mip360_stream_t stream) {
  MIP_REQUIRE(n <= 0 || (x && out), "to8b: null pointer");
  if (n <= 0) return MIP360_OK;
  const int vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  const long long threads = vec ? (n + 15) / 16 : n;
  to8b_kernel<<<(int)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n, out, vec);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

extern "C" int mip360_generate_rays_range(const float* c2w, int c2w_rows, int n_img, int H, int W, float focal,
                                          float near, float far, int ndc, float ndc_near, long long ray_begin,
                                          long long ray_count, float* origins, float* directions, float* viewdirs,
                                          float* radii, float* near_out, float* far_out, mip360_stream_t stream) {
  MIP_REQUIRE(c2w_rows >= 3 && n_img >= 0 && H >= 3 && W >= 3, "generate_rays: bad sizes (need H, W >= 3)");
  const long long total = (long long)n_img * H * W;
  MIP_REQUIRE(ray_begin >= 0 && ray_count >= 0 && ray_begin + ray_count <= total,
              "generate_rays: slab [%lld, %lld) outside the %lld rays of the frame set", ray_begin, ray_begin + ray_count, total);
  if (ray_count == 0) return MIP360_OK;
  MIP_REQUIRE(c2w && origins && directions && viewdirs && radii && near_out && far_out, "generate_rays: null pointer");
  RayGenParams p{c2w, c2w_rows, n_img, H, W, focal, near, far, ndc, ndc_near, ray_begin, ray_count};
  raygen_kernel<<<(int)((ray_count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, origins, directions, viewdirs, radii,
                                                                                 near_out, far_out);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

extern "C" int mip360_generate_rays(const float* c2w, int c2w_rows, int n_img, int H, int W, float focal, float near,
                                    float far, int ndc, float ndc_near, float* origins, float* directions,
                                    float* viewdirs, float* radii, float* near_out, float* far_out,
                                    mip360_stream_t stream) {
  const long long total = (n_img >= 0 && H >= 0 && W >= 0) ? (long long)n_img * H * W : 0;
  return mip360_generate_rays_range(c2w, c2w_rows, n_img, H, W, focal, near, far, ndc, ndc_near, 0, total, origins,
                                    directions, viewdirs, radii, near_out, far_out, stream);
}
