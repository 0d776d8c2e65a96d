// Host-side packing of the SMPL model into the arrays hd_smpl_consts points at (the C counterpart of human_dynamics_b200/smpl.py
// SMPLConstants, so that a C / C++ consumer does not re-implement it).  Pure host code: no CUDA call, usable without a device.
//
// Replaces the data preparation of SMPL.__init__ (src/tf_smpl/batch_smpl.py:27-86): shapedirs (V,3,10) -> (10, V*3) and posedirs
// (V,3,207) -> (207, V*3) row-major (:45-48,60-63), J_regressor pre-composed with v_template / shapedirs (exact refactoring of
// :110-118, evaluated in double), kintree_table[0] cast to int32 with the root's 4294967295 -> -1 (:66), the (V,24) skinning weights as
// ELL (joint ids ascending, zero-weight padding on joint 0), the (K,V) keypoint regressor as CSC over keypoints (:76-82).
#include "common.cuh"

#include <algorithm>
#include <cstdint>
#include <vector>

extern "C" {

int hd_smpl_pack_sizes(int V, int K, const double *weights, const double *kp_regressor, int *lbs_nnz, int *kp_nnz_total) {
  HD_REQUIRE(V > 0 && K > 0 && weights && kp_regressor && lbs_nnz && kp_nnz_total, "hd_smpl_pack_sizes: bad arguments");
  int nnz = 1;
  for (int v = 0; v < V; ++v) {
    int c = 0;
    for (int j = 0; j < 24; ++j) c += weights[(size_t)v * 24 + j] != 0.0;
    nnz = std::max(nnz, c);
  }
  if (nnz <= 4) nnz = 4;
  else if (nnz < 24) nnz = std::min(24, (nnz + 3) / 4 * 4);
  long long tot = 0;
  for (size_t i = 0; i < (size_t)K * V; ++i) tot += kp_regressor[i] != 0.0;
  *lbs_nnz = nnz;
  *kp_nnz_total = (int)tot;
  return HD_OK;
}

int hd_smpl_pack(int V, int K, const double *v_template, const double *shapedirs, const double *posedirs, const double *J_regressor,
                 const double *weights, const double *kp_regressor, const unsigned int *kintree_parents, float *v_template_out, float *dirs,
                 float *J_template, float *J_shapedirs, int *lbs_idx, float *lbs_w, int lbs_nnz, int *kp_ptr, int *kp_vidx, float *kp_w,
                 int *parents) {
  HD_REQUIRE(V > 0 && K > 0 && v_template && shapedirs && posedirs && J_regressor && weights && kp_regressor && kintree_parents &&
                 v_template_out && dirs && J_template && J_shapedirs && lbs_idx && lbs_w && kp_ptr && kp_vidx && kp_w && parents,
             "hd_smpl_pack: null argument");
  int need_nnz = 0, need_tot = 0;
  int r = hd_smpl_pack_sizes(V, K, weights, kp_regressor, &need_nnz, &need_tot);
  if (r) return r;
  HD_REQUIRE(lbs_nnz == need_nnz, "hd_smpl_pack: lbs_nnz must be the value hd_smpl_pack_sizes reports");
  const size_t V3 = (size_t)V * 3;
  for (size_t i = 0; i < V3; ++i) v_template_out[i] = (float)v_template[i];
  // dirs: rows 0..9 = shapedirs reshaped (-1, 10) and transposed, rows 10..216 = posedirs reshaped (-1, 207) and transposed
  for (size_t e = 0; e < V3; ++e) {
    for (int b = 0; b < 10; ++b) dirs[(size_t)b * V3 + e] = (float)shapedirs[e * 10 + b];
    for (int b = 0; b < 207; ++b) dirs[(size_t)(10 + b) * V3 + e] = (float)posedirs[e * 207 + b];
  }
  // J = (beta . shapedirs + v_template) . J_regressor is linear in beta: J_template [24,3], J_shapedirs [10, 24*3]
  for (int j = 0; j < 24; ++j)
    for (int c = 0; c < 3; ++c) {
      double acc = 0.0;
      for (int v = 0; v < V; ++v) acc += J_regressor[(size_t)j * V + v] * v_template[(size_t)v * 3 + c];
      J_template[j * 3 + c] = (float)acc;
      for (int b = 0; b < 10; ++b) {
        double a = 0.0;
        for (int v = 0; v < V; ++v) a += J_regressor[(size_t)j * V + v] * shapedirs[((size_t)v * 3 + c) * 10 + b];
        J_shapedirs[(size_t)b * 72 + j * 3 + c] = (float)a;
      }
    }
  for (int j = 0; j < 24; ++j) parents[j] = (int)(int32_t)kintree_parents[j];            // 4294967295 -> -1
  // ELL skinning weights: the non-zero joints of a vertex in ascending order, then padding (joint 0, weight 0)
  for (int v = 0; v < V; ++v) {
    int n = 0;
    for (int j = 0; j < 24; ++j)
      if (weights[(size_t)v * 24 + j] != 0.0) {
        lbs_idx[(size_t)v * lbs_nnz + n] = j;
        lbs_w[(size_t)v * lbs_nnz + n] = (float)weights[(size_t)v * 24 + j];
        ++n;
      }
    for (; n < lbs_nnz; ++n) { lbs_idx[(size_t)v * lbs_nnz + n] = 0; lbs_w[(size_t)v * lbs_nnz + n] = 0.f; }
  }
  // CSC over keypoints: for keypoint k the vertices with a non-zero coefficient, ascending
  int pos = 0;
  kp_ptr[0] = 0;
  for (int k = 0; k < K; ++k) {
    for (int v = 0; v < V; ++v) {
      const double w = kp_regressor[(size_t)k * V + v];
      if (w != 0.0) { kp_vidx[pos] = v; kp_w[pos] = (float)w; ++pos; }
    }
    kp_ptr[k + 1] = pos;
  }
  return HD_OK;
}

}  // extern "C"

// ---- host bookkeeping of process_image (src/evaluation/run_video.py:69-100 + resize_img, src/util/common.py:7-14): the {Hs, Ws, x0, y0}
// row hd_process_image wants per frame, and the dict entries the reference returns.  Same float64 formulas as the reference's numpy code
// (floor of shape * scale; np.round = round-half-to-even of centre * actual factor, with the reference's own x*fy / y*fx mix-up).
#include <cmath>

extern "C" int hd_crop_geometry(int H, int W, const double *bbox /* cx, cy, scale */, int img_size, int *geom /* Hs, Ws, x0, y0 */,
                                int *center /* 2, nullable */, int *start_pt /* 2, nullable */) {
  HD_REQUIRE(H > 0 && W > 0 && bbox && geom && img_size > 0 && img_size % 2 == 0, "hd_crop_geometry: bad arguments");
  const double scale = bbox[2];
  const long long hs = (long long)std::floor((double)H * scale), ws = (long long)std::floor((double)W * scale);
  HD_REQUIRE(hs >= 1 && ws >= 1 && hs < (1ll << 30) && ws < (1ll << 30), "hd_crop_geometry: bbox scale leaves an empty (or absurd) image");
  const double fy = (double)hs / (double)H, fx = (double)ws / (double)W;
  // center_scaled = np.round(center * [fy, fx]).astype(int)   (run_video.py:75: x is scaled by the y factor and vice versa)
  const long long cx = (long long)std::nearbyint(bbox[0] * fy) + img_size, cy = (long long)std::nearbyint(bbox[1] * fx) + img_size;
  const long long sx = cx - img_size / 2, sy = cy - img_size / 2, ex = cx + img_size / 2, ey = cy + img_size / 2;
  if (sx < 0 || sy < 0 || ex > ws + 2 * img_size || ey > hs + 2 * img_size) {
    hd::set_last_error_text("hd_crop_geometry: bbox centre is more than one crop away from the frame (the reference yields a ragged crop)");
    return HD_ERR_INVALID;
  }
  geom[0] = (int)hs; geom[1] = (int)ws; geom[2] = (int)(sx - img_size); geom[3] = (int)(sy - img_size);
  if (center) { center[0] = (int)(cx - sx); center[1] = (int)(cy - sy); }
  if (start_pt) { start_pt[0] = (int)sx; start_pt[1] = (int)sy; }
  return HD_OK;
}
