/*
 * ORACLE (test infrastructure, not product): plain-C restatement of the reference SMPL path, independent of the numpy
 * one in smpl_ref.py (tests/test_oracle_c.py checks the two against each other).  PARITY: pinned to the
 * reference's source through smpl_ref.py (which is checked against src/tf_smpl/*.py executed over the numpy TensorFlow stand-in,
 * oracle/ref_exec/); unpinned against TensorFlow's own kernels (TF 1.8 cannot run here).
 * Every function follows the cited reference lines; REAL selects float (TF-faithful) or double (truth).
 *
 *   gcc -O2 -shared -fPIC -DREAL=double -o liboracle_smpl_f64.so smpl_ref.c -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL double
#endif

/* src/tf_smpl/batch_lbs.py:42-60 (+ batch_skew :15-39): theta[3] -> R[9], same operation order. */
void oracle_rodrigues(const REAL *theta, REAL *R) {
  const REAL eps = (REAL)1e-8;
  const REAL sx = theta[0] + eps, sy = theta[1] + eps, sz = theta[2] + eps;
  const REAL angle = (REAL)sqrt((double)(sx * sx + sy * sy + sz * sz));      /* tf.norm(theta + 1e-8) */
  const REAL r[3] = {theta[0] / angle, theta[1] / angle, theta[2] / angle};   /* un-shifted theta / angle */
  const REAL c = (REAL)cos((double)angle), s = (REAL)sin((double)angle);
  const REAL skew[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};      /* flat idx [1,2,3,5,6,7] */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      R[i * 3 + j] = c * (i == j ? (REAL)1 : (REAL)0) + ((REAL)1 - c) * (r[i] * r[j]) + s * skew[i * 3 + j];
}

static void mat4_mul(const REAL *a, const REAL *b, REAL *o) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      REAL acc = 0;
      for (int k = 0; k < 4; ++k) acc += a[i * 4 + k] * b[k * 4 + j];
      o[i * 4 + j] = acc;
    }
}

/* src/tf_smpl/batch_lbs.py:133-194 for one sample: Rs[24*9], Js[24*3], parent[24] -> new_J[24*3], A[24*16]. */
void oracle_global_rigid(const REAL *Rs, const REAL *Js, const int *parent, REAL *new_J, REAL *A) {
  REAL res[24][16];
  for (int i = 0; i < 24; ++i) {
    REAL t[3], M[16];
    for (int c = 0; c < 3; ++c) t[c] = (i == 0) ? Js[c] : Js[i * 3 + c] - Js[parent[i] * 3 + c];   /* :170,:173 */
    for (int r = 0; r < 3; ++r) {                                                                 /* make_A :163-168 */
      for (int c = 0; c < 3; ++c) M[r * 4 + c] = Rs[i * 9 + r * 3 + c];
      M[r * 4 + 3] = t[r];
    }
    M[12] = M[13] = M[14] = 0; M[15] = 1;
    if (i == 0) memcpy(res[0], M, sizeof(M));
    else mat4_mul(res[parent[i]], M, res[i]);                                                      /* :175 */
  }
  for (int i = 0; i < 24; ++i) {
    for (int c = 0; c < 3; ++c) new_J[i * 3 + c] = res[i][c * 4 + 3];                               /* :182 */
    memcpy(A + i * 16, res[i], sizeof(REAL) * 16);
    for (int r = 0; r < 4; ++r) {                                                                   /* init_bone :188-192 */
      REAL ib = 0;
      for (int c = 0; c < 3; ++c) ib += res[i][r * 4 + c] * Js[i * 3 + c];
      A[i * 16 + r * 4 + 3] -= ib;
    }
  }
}

/* src/tf_smpl/batch_smpl.py:89-162 for N samples.  Layouts as the reference holds them:
 *   v_template [V*3], shapedirs [10][V*3], posedirs [207][V*3], J_regressor [V][24], weights [V][24], joint_regressor [V][K].
 * Outputs: verts [N][V][3], joints [N][K][3], Rs [N][24][9], Jtr [N][24][3]. */
void oracle_smpl_forward(int N, int V, int K, const REAL *v_template, const REAL *shapedirs, const REAL *posedirs,
                         const REAL *J_regressor, const REAL *weights, const REAL *joint_regressor, const int *parent,
                         const REAL *beta, const REAL *theta, REAL *verts, REAL *joints, REAL *Rs_out, REAL *Jtr) {
  REAL *v_shaped = (REAL *)malloc(sizeof(REAL) * V * 3), *v_posed = (REAL *)malloc(sizeof(REAL) * V * 3);
  for (int n = 0; n < N; ++n) {
    for (int i = 0; i < V * 3; ++i) {                                   /* :110-112 */
      REAL acc = 0;
      for (int b = 0; b < 10; ++b) acc += beta[n * 10 + b] * shapedirs[b * V * 3 + i];
      v_shaped[i] = acc + v_template[i];
    }
    REAL J[72];
    for (int j = 0; j < 24; ++j)                                        /* :115-118 */
      for (int c = 0; c < 3; ++c) {
        REAL acc = 0;
        for (int v = 0; v < V; ++v) acc += v_shaped[v * 3 + c] * J_regressor[v * 24 + j];
        J[j * 3 + c] = acc;
      }
    REAL *Rs = Rs_out + (size_t)n * 216;
    for (int j = 0; j < 24; ++j) oracle_rodrigues(theta + n * 72 + j * 3, Rs + j * 9);      /* :123-124 */
    REAL pf[207];
    for (int q = 0; q < 207; ++q) pf[q] = Rs[9 + q] - ((q % 9 == 0 || q % 9 == 4 || q % 9 == 8) ? (REAL)1 : (REAL)0);   /* :127-128 */
    for (int i = 0; i < V * 3; ++i) {                                   /* :131-133 */
      REAL acc = 0;
      for (int q = 0; q < 207; ++q) acc += pf[q] * posedirs[q * V * 3 + i];
      v_posed[i] = acc + v_shaped[i];
    }
    REAL A[24 * 16];
    oracle_global_rigid(Rs, J, parent, Jtr + (size_t)n * 72, A);        /* :136-137 */
    for (int v = 0; v < V; ++v) {                                       /* :141-151 */
      REAL T[16];
      for (int e = 0; e < 16; ++e) {
        REAL acc = 0;
        for (int k = 0; k < 24; ++k) acc += weights[v * 24 + k] * A[k * 16 + e];
        T[e] = acc;
      }
      for (int r = 0; r < 3; ++r)
        verts[((size_t)n * V + v) * 3 + r] = T[r * 4 + 0] * v_posed[v * 3] + T[r * 4 + 1] * v_posed[v * 3 + 1] +
                                             T[r * 4 + 2] * v_posed[v * 3 + 2] + T[r * 4 + 3] * (REAL)1;
    }
    for (int k = 0; k < K; ++k)                                         /* :154-157 */
      for (int c = 0; c < 3; ++c) {
        REAL acc = 0;
        for (int v = 0; v < V; ++v) acc += verts[((size_t)n * V + v) * 3 + c] * joint_regressor[v * K + k];
        joints[((size_t)n * K + k) * 3 + c] = acc;
      }
  }
  free(v_shaped);
  free(v_posed);
}

/* src/tf_smpl/projection.py:16-29 */
void oracle_orth_proj(int N, int P, const REAL *X, const REAL *cam, REAL *out) {
  for (int n = 0; n < N; ++n)
    for (int p = 0; p < P; ++p) {
      out[((size_t)n * P + p) * 2 + 0] = cam[n * 3] * (X[((size_t)n * P + p) * 3 + 0] + cam[n * 3 + 1]);
      out[((size_t)n * P + p) * 2 + 1] = cam[n * 3] * (X[((size_t)n * P + p) * 3 + 1] + cam[n * 3 + 2]);
    }
}
