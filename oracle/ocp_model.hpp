// TEST INFRASTRUCTURE — CPU restatement (oracle) of the reference's NLP functions.
// Never linked into the product; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build or call it.
//
// Restates, as templated *value* code (derivatives come from ad.hpp):
//   * variable / constraint / parameter layouts      casadi_ocp_formulation.py:88-153, 271-349, 361-376
//   * integration_function (jerk hat functions, t=h)  bound_mpc_functions.py:249-295,
//                                                     jerk_trajectory_casadi.py:78-87,122-131,166-175
//   * fk_pos / velocity_ee / omega_ee                 RobotModel.py:62-100,1055-1107,1270-1303
//     (as the serial-chain product of urdf/body/iiwa14.xacro:65..340; the Maple closed
//      forms are not reproduced — equality to 4e-16 is pinned by tests/golden)
//   * segment selection                               bound_mpc_functions.py:13-20, 34-40
//   * reference_function, bound polynomials           bound_mpc_functions.py:43-149, mpc_utils_casadi.py:140-165
//   * error_function                                  bound_mpc_functions.py:152-202, mpc_utils_casadi.py:6-67
//   * objective_function + sigmoid blend              bound_mpc_functions.py:205-246, casadi_ocp_formulation.py:227-265
//   * inequality rows                                 casadi_ocp_formulation.py:305-349, bound_mpc_functions.py:298-310
// Parity status: pinned against values/derivatives obtained by EXECUTING the reference's
// own Python (tests/golden/make_golden.py); the reference ships no tests of its own.
#pragma once
#include <cmath>
#include "ad.hpp"

namespace orc {

constexpr int NX = 44;   // variables per stage
constexpr int NG = 43;   // constraint rows per stage
constexpr int NE = 36;   // equality rows per stage
constexpr int NI = 7;    // inequality rows per stage (reference form, casadi_ocp_formulation.py:305-349)
constexpr int ND = 12;   // the same feasible set in interval form: rows 36,37 + (m-h<=0, -m-h<=0) for rows 38..42
// offsets inside a stage block  [u(7) u_phi q dq ddq p_pos p_rot v_lin v_ang phi dphi ddphi]
enum { oU = 0, oUPHI = 7, oQ = 8, oDQ = 15, oDDQ = 22, oPPOS = 29, oPROT = 32, oVLIN = 35,
       oVANG = 38, oPHI = 41, oDPHI = 42, oDDPHI = 43 };

struct Layout {  // parameter vector offsets for nr_segs = S
  int S, np;
  int q0 = 0, dq0 = 7, ddq0 = 14, phi0 = 21, p0 = 24, v0 = 30, iwref = 36, dtau = 39, par = 42;
  int orth1, orth2, xphid, jerk, phisw, jacr, jacl, pref, dpref, dpn, bp1, bp2, br1, br2;
  int a4, a3, a2, a1, a0, w, phimax, dphimax, v1, v2, v3, qd;
  explicit Layout(int S_) : S(S_) {
    orth1 = 42 + 3 * S; orth2 = 42 + 6 * S; xphid = 42 + 9 * S; jerk = 45 + 9 * S;
    phisw = 53 + 9 * S; jacr = 54 + 10 * S; jacl = 63 + 10 * S; pref = 72 + 10 * S;
    dpref = 72 + 16 * S; dpn = 72 + 22 * S; bp1 = 72 + 25 * S; bp2 = 72 + 28 * S;
    br1 = 72 + 31 * S; br2 = 72 + 34 * S; a4 = 72 + 37 * S; a3 = 81 + 46 * S;
    a2 = 90 + 55 * S; a1 = 99 + 64 * S; a0 = 108 + 73 * S; w = 117 + 82 * S;
    phimax = 132 + 82 * S; dphimax = 133 + 82 * S; v1 = 134 + 82 * S; v2 = 134 + 85 * S;
    v3 = 134 + 88 * S; qd = 134 + 91 * S; np = 141 + 91 * S;
  }
};

// ---------------------------------------------------------------- kinematics
// joint origin offsets and the constant (signed-permutation) rotations of the chain
static const double JXYZ[7][3] = {{0, 0, 0.1525}, {0, 0, 0.2075}, {0, 0.2325, 0}, {0, 0, 0.1875},
                                  {0, 0.2125, 0}, {0, 0, 0.1875}, {0, 0.0796, 0}};
static const double TOOLZ = 0.2174;
// R = Rz(yaw) Ry(pitch) Rx(roll) of the xacro rpy triples, which are all multiples of pi/2
static const double JROT[7][3][3] = {
    {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}},
    {{-1, 0, 0}, {0, 0, 1}, {0, 1, 0}},
    {{-1, 0, 0}, {0, 0, 1}, {0, 1, 0}},
    {{1, 0, 0}, {0, 0, -1}, {0, 1, 0}},
    {{-1, 0, 0}, {0, 0, 1}, {0, 1, 0}},
    {{1, 0, 0}, {0, 0, -1}, {0, 1, 0}},
    {{-1, 0, 0}, {0, 0, 1}, {0, 1, 0}}};

template <class T>
inline void cross3(const T* a, const T* b, T* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// tool position and tool twist (J(q) dq) of the 7-joint chain
template <class T>
inline void kinematics(const T* q, const T* dq, T* pos, T* vlin, T* vang) {
  T R[3][3], o[3], z[7][3], org[7][3];
  for (int i = 0; i < 3; i++) { o[i] = T(0.0); for (int j = 0; j < 3; j++) R[i][j] = T(i == j ? 1.0 : 0.0); }
  for (int k = 0; k < 7; k++) {
    T Rn[3][3];
    for (int i = 0; i < 3; i++) o[i] = o[i] + R[i][0] * JXYZ[k][0] + R[i][1] * JXYZ[k][1] + R[i][2] * JXYZ[k][2];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        Rn[i][j] = R[i][0] * JROT[k][0][j] + R[i][1] * JROT[k][1][j] + R[i][2] * JROT[k][2][j];
    for (int i = 0; i < 3; i++) { z[k][i] = Rn[i][2]; org[k][i] = o[i]; }
    T c = cos(q[k]), s = sin(q[k]);
    for (int i = 0; i < 3; i++) {
      R[i][0] = Rn[i][0] * c + Rn[i][1] * s;
      R[i][1] = Rn[i][1] * c - Rn[i][0] * s;
      R[i][2] = Rn[i][2];
    }
  }
  for (int i = 0; i < 3; i++) pos[i] = o[i] + R[i][2] * TOOLZ;
  for (int i = 0; i < 3; i++) { vlin[i] = T(0.0); vang[i] = T(0.0); }
  for (int k = 0; k < 7; k++) {
    T r[3], c[3];
    for (int i = 0; i < 3; i++) r[i] = pos[i] - org[k][i];
    cross3(z[k], r, c);
    for (int i = 0; i < 3; i++) { vlin[i] = vlin[i] + c[i] * dq[k]; vang[i] = vang[i] + z[k][i] * dq[k]; }
  }
}

// ---------------------------------------------------------------- segment rules (App. B.1-B.3)
inline int seg_rule1(double phi, const double* phisw, int S) {  // S-row tables
  for (int i = 0; i <= S - 2; i++) if (phi < phisw[i + 1]) return i;
  return S - 1;
}
inline int seg_rule_bp(double phi, const double* phisw, int S) {  // first row of "current and next"
  for (int i = 0; i <= S - 3; i++) if (phi < phisw[i + 1]) return i;
  return S - 2;
}
inline int seg_rule_coef(double phi, const double* phisw, int S) {  // (S+1)-row tables
  for (int i = 0; i <= S - 1; i++) if (phi < phisw[i + 1]) return i;
  return S;
}

// ---------------------------------------------------------------- stage cost (path part) and inequalities
// c = [p_pos(3) p_rot(3) v(6) phi dphi ddphi vprev(6)]  (21 entries)
enum { cPPOS = 0, cPROT = 3, cV = 6, cPHI = 12, cDPHI = 13, cDDPHI = 14, cVPREV = 15, NC = 21 };

template <class T>
inline void path_terms(const Layout& L, const double* p, double dt, const T* c, T& cost, T* ineq, T* din) {
  const int S = L.S;
  const double* phisw = p + L.phisw;
  const T phi = c[cPHI], dphi = c[cDPHI], ddphi = c[cDDPHI];
  const double phiv = valof(phi);
  const int i = seg_rule1(phiv, phisw, S);
  const int jb = seg_rule_bp(phiv, phisw, S);
  const int r = seg_rule_coef(phiv, phisw, S);
  const T tau = phi - phisw[i];
  double dpd[6], dpn[3], bp1[3], bp2[3], br1[3], br2[3], v1[3], v2[3], v3[3], par0[3], o10[3], o20[3];
  T pd[6];
  for (int k = 0; k < 6; k++) { dpd[k] = p[L.dpref + k * S + i]; pd[k] = dpd[k] * tau + p[L.pref + k * S + i]; }
  for (int k = 0; k < 3; k++) {
    dpn[k] = p[L.dpn + k * S + i]; bp1[k] = p[L.bp1 + k * S + jb]; bp2[k] = p[L.bp2 + k * S + jb];
    br1[k] = p[L.br1 + k * S + i]; br2[k] = p[L.br2 + k * S + i];
    v1[k] = p[L.v1 + k * S + i]; v2[k] = p[L.v2 + k * S + i]; v3[k] = p[L.v3 + k * S + i];
    par0[k] = p[L.par + 3 * i + k]; o10[k] = p[L.orth1 + 3 * i + k]; o20[k] = p[L.orth2 + 3 * i + k];
  }
  T b[9];
  for (int j = 0; j < 9; j++) {
    const int o = j * (S + 1) + r;
    T t2 = tau * tau;
    b[j] = p[L.a4 + o] * (t2 * t2) + p[L.a3 + o] * (t2 * tau) + p[L.a2 + o] * t2 + p[L.a1 + o] * tau + p[L.a0 + o];
  }
  // position error (mpc_utils_casadi.py:19-67)
  T ep[3], eppar[3], tdot = T(0.0);
  for (int k = 0; k < 3; k++) { ep[k] = c[cPPOS + k] - pd[k]; tdot = tdot + dpd[k] * ep[k]; }
  for (int k = 0; k < 3; k++) eppar[k] = tdot * dpd[k];
  // orientation error (mpc_utils_casadi.py:6-10), jacobians stored column-major
  T dl[3], dr[3], er[3], d[3];
  for (int k = 0; k < 3; k++) { dl[k] = c[cPROT + k] - p[L.p0 + 3 + k]; dr[k] = pd[3 + k] - p[L.iwref + k]; }
  for (int k = 0; k < 3; k++) {
    d[k] = T(0.0);
    for (int m = 0; m < 3; m++) d[k] = d[k] + p[L.jacl + m * 3 + k] * dl[m] - p[L.jacr + m * 3 + k] * dr[m];
    er[k] = p[L.dtau + k] + d[k];
  }
  T s1 = T(0.0), sp = T(0.0), s2 = T(0.0);
  for (int k = 0; k < 3; k++) { s1 = s1 + d[k] * v1[k]; sp = sp + d[k] * v2[k]; s2 = s2 + d[k] * v3[k]; }
  T ero1[3], erpar[3], ero2[3];
  for (int k = 0; k < 3; k++) { ero1[k] = o10[k] + s1 * br1[k]; erpar[k] = par0[k] + sp * dpn[k]; ero2[k] = o20[k] + s2 * br2[k]; }
  // objective (casadi_ocp_formulation.py:227-265, bound_mpc_functions.py:205-246)
  const double* w = p + L.w;
  T sig = 1.0 / (1.0 + exp(-100.0 * (phi - (p[L.phimax] - 0.02))));
  cost = T(0.0);
  for (int k = 0; k < 3; k++) {
    T eo = sig * er[k] + (1.0 - sig) * erpar[k];
    cost = cost + w[1] * (eo * eo);
  }
  for (int k = 0; k < 3; k++) {
    T eo = sig * ep[k] + (1.0 - sig) * eppar[k];
    cost = cost + w[0] * (eo * eo);
  }
  for (int k = 0; k < 6; k++) {
    T dv = c[cV + k] - dphi * dpd[k];
    cost = cost + w[2] * (dv * dv);
    T da = (c[cV + k] - c[cVPREV + k]) / dt - ddphi * dpd[k];
    cost = cost + w[5] * (da * da);
  }
  T e0 = p[L.xphid] - phi, e1 = p[L.xphid + 1] - dphi, e2 = p[L.xphid + 2] - ddphi;
  cost = cost + w[6] * (e0 * e0) + w[7] * (e1 * e1) + w[8] * (e2 * e2);
  // inequalities (casadi_ocp_formulation.py:305-349)
  ineq[0] = phi - p[L.phimax];
  ineq[1] = dphi - p[L.dphimax];
  T proj = T(0.0);
  for (int k = 0; k < 3; k++) proj = proj + dpn[k] * erpar[k];
  ineq[2] = proj * proj - b[8] * b[8];
  T e1p = T(0.0), e2p = T(0.0), r1p = T(0.0), r2p = T(0.0);
  for (int k = 0; k < 3; k++) {
    e1p = e1p + ep[k] * bp1[k]; e2p = e2p + ep[k] * bp2[k];
    r1p = r1p + br1[k] * ero1[k]; r2p = r2p + br2[k] * ero2[k];
  }
  T m0 = e1p - 0.5 * (b[0] + b[2]), m1 = e2p - 0.5 * (b[1] + b[3]);
  T h0 = (b[0] - b[2]) / 2.0, h1 = (b[1] - b[3]) / 2.0;
  ineq[3] = m0 * m0 - h0 * h0;
  ineq[4] = m1 * m1 - h1 * h1;
  T m2 = r1p - 0.5 * (b[4] + b[6]), m3 = r2p - 0.5 * (b[5] + b[7]);
  T h2 = (b[4] - b[6]) / 2.0, h3 = (b[5] - b[7]) / 2.0;
  ineq[5] = m2 * m2 - h2 * h2;
  ineq[6] = m3 * m3 - h3 * h3;
  // interval form used inside the interior-point iteration: (m)^2 - h^2 <= 0  <=>  -h <= m <= h  (h > 0)
  din[0] = ineq[0]; din[1] = ineq[1];
  din[2] = proj - b[8]; din[3] = -proj - b[8];
  din[4] = m0 - h0; din[5] = -m0 - h0;
  din[6] = m1 - h1; din[7] = -m1 - h1;
  din[8] = m2 - h2; din[9] = -m2 - h2;
  din[10] = m3 - h3; din[11] = -m3 - h3;
}

// integration coefficients of the piecewise-linear jerk at t = h
struct IntCoef {
  double h, a_dq, a_ddq, a_um, a_u, b_ddq, b_um, b_u, c_um, c_u;
  explicit IntCoef(double h_) : h(h_) {
    a_dq = h; a_ddq = h * h / 2; a_um = h * h * h / 8; a_u = h * h * h / 24;
    b_ddq = h; b_um = h * h / 3; b_u = h * h / 6; c_um = h / 2; c_u = h / 2;
  }
};

}  // namespace orc
