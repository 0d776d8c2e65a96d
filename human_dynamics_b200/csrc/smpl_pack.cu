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
