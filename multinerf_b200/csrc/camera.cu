// Pixel -> ray generation on the device: one thread per ray.
//
// Replaces (reference file:line): camera_utils.pixels_to_rays camera_utils.py:522-636,
// _compute_residual_and_jacobian :427-475, _radial_and_tangential_undistort :478-513,
// convert_to_ndc :32-97 and the per-ray camera gather of cast_ray_batch :639-688 -- the
// `xnp=jnp` path of train_utils.py:266-268 (Config.cast_rays_in_train_step).
// HBM-bound: reads 12 B (pixel + camera index; the camera matrices stay in L1/L2), writes 48 B
// per ray.  Compiled without FMA contraction so the fp32 rounding follows the reference's
// unfused elementwise graph.
#include "common.cuh"

namespace mnrf {

struct V3 { float x, y, z; };

__device__ __forceinline__ V3 mat3_vec(const float* __restrict__ m, int ld, V3 v) {
  V3 r;
  r.x = m[0] * v.x + m[1] * v.y + m[2] * v.z;
  r.y = m[ld] * v.x + m[ld + 1] * v.y + m[ld + 2] * v.z;
  r.z = m[2 * ld] * v.x + m[2 * ld + 1] * v.y + m[2 * ld + 2] * v.z;
  return r;
}

__device__ __forceinline__ void undistort(const mnrf_camera_desc& d, float xd, float yd, float& xo, float& yo) {
  float x = xd, y = yd;
  const float k1 = d.k1, k2 = d.k2, k3 = d.k3, k4 = d.k4, p1 = d.p1, p2 = d.p2;
  for (int it = 0; it < d.undistort_iters; ++it) {
    const float r = x * x + y * y;
    const float dd = 1.0f + r * (k1 + r * (k2 + r * (k3 + r * k4)));
    const float fx = dd * x + 2.f * p1 * x * y + p2 * (r + 2.f * x * x) - xd;
    const float fy = dd * y + 2.f * p2 * x * y + p1 * (r + 2.f * y * y) - yd;
    const float d_r = k1 + r * (2.0f * k2 + r * (3.0f * k3 + r * 4.0f * k4));
    const float d_x = 2.0f * x * d_r;
    const float d_y = 2.0f * y * d_r;
    const float fx_x = dd + d_x * x + 2.0f * p1 * y + 6.0f * p2 * x;
    const float fx_y = d_y * x + 2.0f * p1 * x + 2.0f * p2 * y;
    const float fy_x = d_x * y + 2.0f * p2 * y + 2.0f * p1 * x;
    const float fy_y = dd + d_y * y + 2.0f * p2 * x + 6.0f * p1 * y;
    const float den = fy_x * fx_y - fx_x * fy_y;
    const float xn = fx * fy_y - fy * fx_y;
    const float yn = fy * fx_x - fx * fy_x;
    const bool ok = fabsf(den) > d.undistort_eps;
    x = x + (ok ? xn / den : 0.f);
    y = y + (ok ? yn / den : 0.f);
  }
  xo = x; yo = y;
}

// camera-space direction of pixel centre (px, py): inverse intrinsics, undistortion, fisheye,
// OpenCV -> OpenGL flip
__device__ __forceinline__ V3 camera_dir(const mnrf_camera_desc& d, const float* __restrict__ p2c, float px, float py) {
  V3 v = mat3_vec(p2c, 3, V3{px + 0.5f, py + 0.5f, 1.0f});
  if (d.has_distortion) {
    float x, y;
    undistort(d, v.x, v.y, x, y);
    v = V3{x, y, 1.0f};
  }
  if (d.camtype == MNRF_CAM_FISHEYE) {
    float theta = sqrtf(v.x * v.x + v.y * v.y);
    theta = fminf(3.14159274101257324f, theta);
    const float s = sinf(theta) / theta;
    v = V3{v.x * s, v.y * s, cosf(theta)};
  }
  return V3{v.x, -v.y, -v.z};
}

// convert_to_ndc: returns the NDC origin; `dir` is overwritten with the NDC direction
__device__ __forceinline__ V3 to_ndc(const mnrf_camera_desc& d, V3 o, V3& dir) {
  const float t = -(d.ndc_near + o.z) / dir.z;
  o = V3{o.x + t * dir.x, o.y + t * dir.y, o.z + t * dir.z};
  const float xm = 1.0f / d.ndc_p02, ym = 1.0f / d.ndc_p12;
  const V3 o_ndc{xm * o.x / o.z, ym * o.y / o.z, -1.0f};
  const V3 inf_ndc{xm * dir.x / dir.z, ym * dir.y / dir.z, 1.0f};
  dir = V3{inf_ndc.x - o_ndc.x, inf_ndc.y - o_ndc.y, inf_ndc.z - o_ndc.z};
  return o_ndc;
}

__device__ __forceinline__ float dist3(V3 a, V3 b) {
  const float x = a.x - b.x, y = a.y - b.y, z = a.z - b.z;
  return sqrtf(x * x + y * y + z * z);
}

__global__ void __launch_bounds__(256)
pixels_to_rays_kernel(mnrf_camera_desc d, const int32_t* __restrict__ pix_x, const int32_t* __restrict__ pix_y,
                      const int32_t* __restrict__ cam_idx, const float* __restrict__ pixtocams,
                      const float* __restrict__ camtoworlds, float* __restrict__ origins,
                      float* __restrict__ directions, float* __restrict__ viewdirs,
                      float* __restrict__ radii, float* __restrict__ imageplane) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.num_rays; i += gridDim.x * blockDim.x) {
    int cam = (d.num_cameras > 1 && cam_idx) ? cam_idx[i] : 0;
    cam = min(max(cam, 0), d.num_cameras - 1);
    const float* p2c = pixtocams + (size_t)cam * 9;
    const float* c2w = camtoworlds + (size_t)cam * 12;
    const int xi = pix_x[i], yi = pix_y[i];
    const V3 c0 = camera_dir(d, p2c, (float)xi, (float)yi);
    const V3 cx = camera_dir(d, p2c, (float)(xi + 1), (float)yi);
    const V3 cy = camera_dir(d, p2c, (float)xi, (float)(yi + 1));
    V3 dir = mat3_vec(c2w, 4, c0);
    V3 dx = mat3_vec(c2w, 4, cx);
    V3 dy = mat3_vec(c2w, 4, cy);
    V3 o{c2w[3], c2w[7], c2w[11]};
    const float n = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    const V3 vd{dir.x / n, dir.y / n, dir.z / n};
    float dx_norm, dy_norm;
    if (!d.has_ndc) {
      dx_norm = dist3(dx, dir);
      dy_norm = dist3(dy, dir);
    } else {
      const V3 o_dx = to_ndc(d, o, dx);
      const V3 o_dy = to_ndc(d, o, dy);
      o = to_ndc(d, o, dir);
      dx_norm = dist3(o_dx, o);
      dy_norm = dist3(o_dy, o);
    }
    origins[3 * i + 0] = o.x; origins[3 * i + 1] = o.y; origins[3 * i + 2] = o.z;
    directions[3 * i + 0] = dir.x; directions[3 * i + 1] = dir.y; directions[3 * i + 2] = dir.z;
    viewdirs[3 * i + 0] = vd.x; viewdirs[3 * i + 1] = vd.y; viewdirs[3 * i + 2] = vd.z;
    radii[i] = (0.5f * (dx_norm + dy_norm)) * 2.f / 3.4641016151377544f;
    imageplane[2 * i + 0] = c0.x; imageplane[2 * i + 1] = c0.y;
  }
}

}  // namespace mnrf

extern "C" int mnrf_pixels_to_rays(const mnrf_camera_desc* d, const int32_t* pix_x, const int32_t* pix_y,
                                   const int32_t* cam_idx, const float* pixtocams, const float* camtoworlds,
                                   float* origins, float* directions, float* viewdirs, float* radii,
                                   float* imageplane, mnrf_stream stream) {
  using namespace mnrf;
  set_error("");
  if (d && d->num_rays == 0) return 0;            // nothing to do (and empty tensors carry null pointers)
  MNRF_CHECK(d && pix_x && pix_y && pixtocams && camtoworlds && origins && directions && viewdirs && radii &&
             imageplane, "mnrf_pixels_to_rays: null pointer");
  MNRF_CHECK(d->num_cameras >= 1, "mnrf_pixels_to_rays: num_cameras must be >= 1");
  MNRF_CHECK(d->num_cameras == 1 || cam_idx, "mnrf_pixels_to_rays: cam_idx is required with several cameras");
  MNRF_CHECK(d->camtype == MNRF_CAM_PERSPECTIVE || d->camtype == MNRF_CAM_FISHEYE,
             "mnrf_pixels_to_rays: camtype must be perspective or fisheye");
  MNRF_CHECK(!d->has_distortion || d->undistort_iters >= 0, "mnrf_pixels_to_rays: undistort_iters < 0");
  if (d->num_rays == 0) return 0;
  int blocks = ceil_div(d->num_rays, 256);
  const int maxb = mnrf_num_sms() * 8;
  if (blocks > maxb) blocks = maxb;
  pixels_to_rays_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*d, pix_x, pix_y, cam_idx, pixtocams,
                                                                  camtoworlds, origins, directions, viewdirs,
                                                                  radii, imageplane);
  MNRF_LAUNCH_CHECK();
  return 0;
}
