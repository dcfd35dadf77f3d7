// HalfCheetah physics, ONE THREAD PER ENV variant (the default; see DESIGN.md for the
// measurement that made it so: the warp-per-env kernel spends 34k warp-instructions per env
// step with ~16 of 32 lanes active, this one ~2.5k warp-instructions for 32 envs' worth).
//
// Same pipeline as mujoco.cu's warp kernel and as MuJoCo documents it, specialised for the
// planar cheetah so that everything but the constraint-row table lives in registers:
//   * composite-rigid-body inertia in planar form (mj_crb): M[i][j] = I_C + m_C r_i.r_j
//   * the joint-space matrices M, H = M + J^T D J and M + h B all share one block-arrow
//     sparsity -- root dofs {0,1,2} couple to both legs, back-leg dofs {3,4,5} and front-leg
//     dofs {6,7,8} never couple (no constraint row touches both legs) -- so they are stored
//     as 36 doubles (R 3x3 sym, A_b, A_f 3x3 sym, C_b, C_f 3x3) and factorised leaves-first,
//     which is what MuJoCo's sparse L^T D L does on the kinematic tree
//   * bias forces by a planar Newton-Euler pass with subtree force/moment sums (mj_rne)
//   * constraint rows stored sparse: 3 root entries + 3 leg entries + leg id
#pragma once

namespace epb {
namespace hct {

constexpr int NV = 9, NB = 7, NG = 8, NU = 6;
constexpr int MAXROW = 6 + 3 * 16;

// packed symmetric 3x3: [0]=(0,0) [1]=(1,0) [2]=(1,1) [3]=(2,0) [4]=(2,1) [5]=(2,2)
struct Arrow {
  double R[6], Ab[6], Af[6];
  double Cb[3][3], Cf[3][3];  // [leg dof][root dof]
};

// y = A x for packed symmetric 3x3
__device__ __forceinline__ void symv3(const double* A, const double* x, double* y) {
  y[0] = A[0] * x[0] + A[1] * x[1] + A[3] * x[2];
  y[1] = A[1] * x[0] + A[2] * x[1] + A[4] * x[2];
  y[2] = A[3] * x[0] + A[4] * x[1] + A[5] * x[2];
}

// y = H x for the block-arrow matrix; x, y are [root(3) | back(3) | front(3)]
__device__ __forceinline__ void arrow_mv(const Arrow& H, const double* x, double* y) {
  symv3(H.R, x, y);
  symv3(H.Ab, x + 3, y + 3);
  symv3(H.Af, x + 6, y + 6);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    y[r] += H.Cb[0][r] * x[3] + H.Cb[1][r] * x[4] + H.Cb[2][r] * x[5] +
            H.Cf[0][r] * x[6] + H.Cf[1][r] * x[7] + H.Cf[2][r] * x[8];
  }
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    y[3 + l] += H.Cb[l][0] * x[0] + H.Cb[l][1] * x[1] + H.Cb[l][2] * x[2];
    y[6 + l] += H.Cf[l][0] * x[0] + H.Cf[l][1] * x[1] + H.Cf[l][2] * x[2];
  }
}

// in-place Cholesky of a packed symmetric 3x3: A = L L^T; the diagonal of L is stored
// INVERTED (multiplications instead of divisions in the substitutions)
__device__ __forceinline__ void chol3(double* A) {
  double d0 = rsqrt(A[0]);
  double l10 = A[1] * d0, l20 = A[3] * d0;
  double d1 = rsqrt(A[2] - l10 * l10);
  double l21 = (A[4] - l20 * l10) * d1;
  double d2 = rsqrt(A[5] - l20 * l20 - l21 * l21);
  A[0] = d0; A[1] = l10; A[2] = d1; A[3] = l20; A[4] = l21; A[5] = d2;
}
// x <- L^-1 x
__device__ __forceinline__ void fwd3(const double* L, double* x) {
  x[0] = x[0] * L[0];
  x[1] = (x[1] - L[1] * x[0]) * L[2];
  x[2] = (x[2] - L[3] * x[0] - L[4] * x[1]) * L[5];
}
// x <- L^-T x
__device__ __forceinline__ void bwd3(const double* L, double* x) {
  x[2] = x[2] * L[5];
  x[1] = (x[1] - L[4] * x[2]) * L[2];
  x[0] = (x[0] - L[1] * x[1] - L[3] * x[2]) * L[0];
}

// Solve H x = g (x overwrites g).  H is destroyed.  Leaves-first block elimination:
//   A_l = L_l L_l^T,  W_l = L_l^-1 C_l,  S = R - W_b^T W_b - W_f^T W_f = L_s L_s^T
__device__ __noinline__ void arrow_solve(Arrow& H, double* g) {
  chol3(H.Ab);
  chol3(H.Af);
  // W = L^-1 C, column by column (each root dof r is one right-hand side)
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double cb[3] = {H.Cb[0][r], H.Cb[1][r], H.Cb[2][r]};
    double cf[3] = {H.Cf[0][r], H.Cf[1][r], H.Cf[2][r]};
    fwd3(H.Ab, cb);
    fwd3(H.Af, cf);
    H.Cb[0][r] = cb[0]; H.Cb[1][r] = cb[1]; H.Cb[2][r] = cb[2];
    H.Cf[0][r] = cf[0]; H.Cf[1][r] = cf[1]; H.Cf[2][r] = cf[2];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double s = 0;
#pragma unroll
      for (int l = 0; l < 3; ++l) s += H.Cb[l][i] * H.Cb[l][j] + H.Cf[l][i] * H.Cf[l][j];
      H.R[i * (i + 1) / 2 + j] -= s;
    }
  chol3(H.R);
  fwd3(H.Ab, g + 3);
  fwd3(H.Af, g + 6);
#pragma unroll
  for (int r = 0; r < 3; ++r)
    g[r] -= H.Cb[0][r] * g[3] + H.Cb[1][r] * g[4] + H.Cb[2][r] * g[5] +
            H.Cf[0][r] * g[6] + H.Cf[1][r] * g[7] + H.Cf[2][r] * g[8];
  fwd3(H.R, g);
  bwd3(H.R, g);
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    g[3 + l] -= H.Cb[l][0] * g[0] + H.Cb[l][1] * g[1] + H.Cb[l][2] * g[2];
    g[6 + l] -= H.Cf[l][0] * g[0] + H.Cf[l][1] * g[1] + H.Cf[l][2] * g[2];
  }
  bwd3(H.Ab, g + 3);
  bwd3(H.Af, g + 6);
}

// sparse constraint row: J = [jr(3 root dofs) | jl(3 dofs of one leg)]
struct Rows {
  double jr[MAXROW][3], jl[MAXROW][3];
  double D[MAXROW], aref[MAXROW], jar[MAXROW], Jv[MAXROW];
  signed char leg[MAXROW];  // 0 = back (dofs 3..5), 1 = front (dofs 6..8), -1 = none
  int n;
};

__device__ __forceinline__ double row_dot(const Rows& e, int r, const double* x) {
  double s = e.jr[r][0] * x[0] + e.jr[r][1] * x[1] + e.jr[r][2] * x[2];
  int l = e.leg[r];
  if (l >= 0) {
    const double* xl = x + 3 + 3 * l;
    s += e.jl[r][0] * xl[0] + e.jl[r][1] * xl[1] + e.jl[r][2] * xl[2];
  }
  return s;
}

__device__ __noinline__ void sincos_ool(double x, double* s, double* c) { sincos(x, s, c); }

struct HcState {
  double q[NV], v[NV], warm[NV], ctrl[NU];
};

// one mj_step for one env, everything thread-private
__device__ void substep(HcState& s, Rows& e) {
  // ---- kinematics (body order: torso, bthigh, bshin, bfoot, fthigh, fshin, ffoot) ----
  double ox[NB], oz[NB], c[NB], sn[NB], cx[NB], cz[NB], om[NB], aox[NB], aoz[NB];
  {
    double th[NB];
    th[0] = s.q[2];
    th[1] = th[0] + s.q[3]; th[2] = th[1] + s.q[4]; th[3] = th[2] + s.q[5];
    th[4] = th[0] + s.q[6]; th[5] = th[4] + s.q[7]; th[6] = th[5] + s.q[8];
    om[0] = s.v[2];
    om[1] = om[0] + s.v[3]; om[2] = om[1] + s.v[4]; om[3] = om[2] + s.v[5];
    om[4] = om[0] + s.v[6]; om[5] = om[4] + s.v[7]; om[6] = om[5] + s.v[8];
    // one out-of-line copy of the double-precision sincos instead of seven inlined ones:
    // the substep is ~80 KB of SASS and instruction fetch is a measured stall at 7 warps/SM
#pragma unroll
    for (int b = 0; b < NB; ++b) sincos_ool(th[b], &sn[b], &c[b]);
  }
  ox[0] = cm.bposx[0] + s.q[0];
  oz[0] = cm.bposz[0] + s.q[1];
  aox[0] = 0; aoz[0] = 0;
#pragma unroll
  for (int b = 1; b < NB; ++b) {
    const int p = (b == 4) ? 0 : b - 1;  // parents {-1,0,1,2,0,4,5}
    double rx = c[p] * cm.bposx[b] + sn[p] * cm.bposz[b];
    double rz = -sn[p] * cm.bposx[b] + c[p] * cm.bposz[b];
    ox[b] = ox[p] + rx;
    oz[b] = oz[p] + rz;
    double op2 = om[p] * om[p];
    aox[b] = aox[p] - op2 * rx;
    aoz[b] = aoz[p] - op2 * rz;
  }
  // CoM, inertial force per body (m (a - g)) and its moment about the world origin
  double fx[NB], fz[NB], tq[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    double rx = c[b] * cm.comx[b] + sn[b] * cm.comz[b];
    double rz = -sn[b] * cm.comx[b] + c[b] * cm.comz[b];
    cx[b] = ox[b] + rx;
    cz[b] = oz[b] + rz;
    double o2 = om[b] * om[b];
    fx[b] = cm.mass[b] * (aox[b] - o2 * rx);
    fz[b] = cm.mass[b] * (aoz[b] - o2 * rz - cm.gravity);
    tq[b] = cz[b] * fx[b] - cx[b] * fz[b];
  }
  // ---- composite bodies, leaves to root: (mass, com, inertia about com), force sums ----
  double Cm[NB], Ccx[NB], Ccz[NB], Ci[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    Cm[b] = cm.mass[b]; Ccx[b] = cx[b]; Ccz[b] = cz[b]; Ci[b] = cm.iyy[b];
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    // children merged into parents in the order foot->shin->thigh->torso for both legs
    const int child = (k == 0) ? 3 : (k == 1) ? 2 : (k == 2) ? 1 : (k == 3) ? 6 : (k == 4) ? 5 : 4;
    const int par = (child == 4 || child == 1) ? 0 : child - 1;
    double m = Cm[par] + Cm[child];
    double minv = 1.0 / m;
    double nx = (Cm[par] * Ccx[par] + Cm[child] * Ccx[child]) * minv;
    double nz = (Cm[par] * Ccz[par] + Cm[child] * Ccz[child]) * minv;
    double dpx = Ccx[par] - nx, dpz = Ccz[par] - nz, dcx = Ccx[child] - nx, dcz = Ccz[child] - nz;
    Ci[par] = Ci[par] + Ci[child] + Cm[par] * (dpx * dpx + dpz * dpz) +
              Cm[child] * (dcx * dcx + dcz * dcz);
    Cm[par] = m; Ccx[par] = nx; Ccz[par] = nz;
    fx[par] += fx[child]; fz[par] += fz[child]; tq[par] += tq[child];
  }
  // ---- joint-space inertia (block arrow) + armature ------------------------------------
  Arrow M;
  M.R[0] = Cm[0]; M.R[1] = 0; M.R[2] = Cm[0];
  {
    double rx = Ccx[0] - ox[0], rz = Ccz[0] - oz[0];
    M.R[3] = Cm[0] * rz;
    M.R[4] = -Cm[0] * rx;
    M.R[5] = Ci[0] + Cm[0] * (rx * rx + rz * rz) + cm.armature[2];
  }
#pragma unroll
  for (int leg = 0; leg < 2; ++leg) {
    double* A = leg ? M.Af : M.Ab;
    double (*C)[3] = leg ? M.Cf : M.Cb;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      const int b = 1 + 3 * leg + l;  // body whose hinge is leg dof l
      double rx = Ccx[b] - ox[b], rz = Ccz[b] - oz[b];
      C[l][0] = Cm[b] * rz;
      C[l][1] = -Cm[b] * rx;
      C[l][2] = Ci[b] + Cm[b] * (rx * (Ccx[b] - ox[0]) + rz * (Ccz[b] - oz[0]));
#pragma unroll
      for (int j = 0; j <= l; ++j) {
        const int a = 1 + 3 * leg + j;  // ancestor-or-self hinge body
        double v = Ci[b] + Cm[b] * (rx * (Ccx[b] - ox[a]) + rz * (Ccz[b] - oz[a]));
        if (j == l) v += cm.armature[3 + 3 * leg + l];
        A[l * (l + 1) / 2 + j] = v;
      }
    }
  }
  // ---- bias (subtree moment about each hinge), passive, actuation -> qfrc_smooth --------
  double fs[NV];
  fs[0] = -fx[0];
  fs[1] = -fz[0];
  fs[2] = -(tq[0] - oz[0] * fx[0] + ox[0] * fz[0]);
#pragma unroll
  for (int i = 3; i < NV; ++i) {
    const int b = i - 2;
    double bias = tq[b] - oz[b] * fx[b] + ox[b] * fz[b];
    double ctrl = s.ctrl[i - 3];
    ctrl = ctrl < -1 ? -1 : (ctrl > 1 ? 1 : ctrl);
    fs[i] = -cm.stiffness[i] * s.q[i] - cm.damping[i] * s.v[i] - bias + cm.gear[i - 3] * ctrl;
  }
  double as[NV];
  {
    Arrow L = M;
#pragma unroll
    for (int i = 0; i < NV; ++i) as[i] = fs[i];
    arrow_solve(L, as);  // qacc_smooth
  }
  // ---- collision + constraint rows ---------------------------------------------------------
  // Two phases so that the expensive row construction is not replicated (and predicated
  // off for almost every lane) once per candidate site: first collect the few active limits
  // / contacts of this env into a list, then build rows in a loop over that list.
  e.n = 0;
  int nlim = 0, ncon = 0;
  signed char lim_j[6];
  double lim_dist[6];
  signed char con_b[16];
  double con_px[16], con_pz[16];
#pragma unroll
  for (int j = 0; j < 6; ++j) {  // joint limits (mj_instantiateLimit)
    const int i = 3 + j;
    double dlo = s.q[i] - cm.rlo[i], dhi = cm.rhi[i] - s.q[i];
    if (dlo < 0) {
      lim_j[nlim] = (signed char)j; lim_dist[nlim] = dlo; ++nlim;
    } else if (dhi < 0) {
      lim_j[nlim] = (signed char)(j + 8); lim_dist[nlim] = dhi; ++nlim;  // +8 tags the upper side
    }
  }
#pragma unroll
  for (int g = 0; g < NG; ++g) {  // floor plane vs capsule end spheres, margin 0
    const int b = (g <= 1) ? 0 : g - 1;
    double gx = ox[b] + c[b] * cm.gposx[g] + sn[b] * cm.gposz[g];
    double gz = oz[b] - sn[b] * cm.gposx[g] + c[b] * cm.gposz[g];
    double ax = c[b] * cm.gaxx[g] + sn[b] * cm.gaxz[g];
    double az = -sn[b] * cm.gaxx[g] + c[b] * cm.gaxz[g];
#pragma unroll
    for (int en = 0; en < 2; ++en) {
      double sg = en ? -1.0 : 1.0;
      double pz = gz + sg * cm.ghalf[g] * az;
      if (!(pz > cm.grad)) {
        con_b[ncon] = (signed char)b;
        con_px[ncon] = gx + sg * cm.ghalf[g] * ax;
        con_pz[ncon] = pz;
        ++ncon;
      }
    }
  }
  for (int k = 0; k < nlim; ++k) {
    const int j = lim_j[k] & 7;
    const double sign = (lim_j[k] & 8) ? -1.0 : 1.0, dist = lim_dist[k];
    const int i = 3 + j;
    int r = e.n++;
    double imp, K, B;
    impedance(cm.solref_limit, cm.solimp_limit, dist, imp, K, B);
    double Rr = fmax(MINVAL, (1 - imp) * cm.dof_invweight0[i] / imp);
    e.jr[r][0] = e.jr[r][1] = e.jr[r][2] = 0;
    e.jl[r][0] = (j % 3 == 0) ? sign : 0.0;
    e.jl[r][1] = (j % 3 == 1) ? sign : 0.0;
    e.jl[r][2] = (j % 3 == 2) ? sign : 0.0;
    e.leg[r] = (signed char)(j / 3);
    e.D[r] = 1 / Rr;
    e.aref[r] = -B * (sign * s.v[i]) - K * imp * dist;
  }
  for (int k = 0; k < ncon; ++k) {
    const int b = con_b[k];
    const double px = con_px[k], pz = con_pz[k];
    const double dist = pz - cm.grad;
    const double cpz = pz - (cm.grad + dist / 2);  // sphere centre - n (radius + dist/2)
    // point Jacobian: root dofs, then the hinges of the body's own leg down to its level
    const int leg = (b == 0) ? -1 : (b - 1) / 3, lvl = (b == 0) ? -1 : (b - 1) % 3;
    const bool fr = leg == 1;
    const double hx[3] = {fr ? ox[4] : ox[1], fr ? ox[5] : ox[2], fr ? ox[6] : ox[3]};
    const double hz[3] = {fr ? oz[4] : oz[1], fr ? oz[5] : oz[2], fr ? oz[6] : oz[3]};
    const double jx[3] = {1, 0, cpz - oz[0]}, jz[3] = {0, 1, -(px - ox[0])};
    double lx[3], lz[3];
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      lx[l] = (l <= lvl) ? cpz - hz[l] : 0.0;
      lz[l] = (l <= lvl) ? -(px - hx[l]) : 0.0;
    }
    double velx = jx[0] * s.v[0] + jx[2] * s.v[2];
    double velz = jz[1] * s.v[1] + jz[2] * s.v[2];
    {
      const double v0 = fr ? s.v[6] : s.v[3], v1 = fr ? s.v[7] : s.v[4], v2 = fr ? s.v[8] : s.v[5];
      velx += lx[0] * v0 + lx[1] * v1 + lx[2] * v2;
      velz += lz[0] * v0 + lz[1] * v1 + lz[2] * v2;
    }
    double imp, K, B;
    impedance(cm.solref, cm.solimp, dist, imp, K, B);
    const double tran = cm.body_invw_tran[b];
    const double dA = tran + cm.mu * cm.mu * tran;
    const double Rr = fmax(MINVAL, (1 - imp) * dA / imp) * (2 * cm.mu * cm.mu);
    const double D = 1 / Rr, kip = K * imp * dist;
    const int r = e.n;
    e.n += 3;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      e.jr[r][q] = jz[q] + cm.mu * jx[q];
      e.jr[r + 1][q] = jz[q] - cm.mu * jx[q];
      e.jr[r + 2][q] = jz[q];
      e.jl[r][q] = lz[q] + cm.mu * lx[q];
      e.jl[r + 1][q] = lz[q] - cm.mu * lx[q];
      e.jl[r + 2][q] = lz[q];
    }
    e.leg[r] = e.leg[r + 1] = e.leg[r + 2] = (signed char)leg;
    e.D[r] = D; e.D[r + 1] = D; e.D[r + 2] = 2 * D;
    e.aref[r] = -B * (velz + cm.mu * velx) - kip;
    e.aref[r + 1] = -B * (velz - cm.mu * velx) - kip;
    e.aref[r + 2] = -B * velz - kip;
  }
  // ---- constraint solve (Newton, exact line search) -------------------------------------------
  double a[NV], fc[NV];
  const int n = e.n;
  if (n == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) { a[i] = as[i]; fc[i] = 0; }
  } else {
    double Ma[NV], grad[NV], srch[NV], Mv[NV];
    {  // warmstart choice
      double cw = 0, cs = 0;
      for (int r = 0; r < n; ++r) {
        double sw = row_dot(e, r, s.warm) - e.aref[r];
        double ss = row_dot(e, r, as) - e.aref[r];
        if (sw < 0) cw += 0.5 * e.D[r] * sw * sw;
        if (ss < 0) cs += 0.5 * e.D[r] * ss * ss;
      }
      arrow_mv(M, s.warm, Ma);
#pragma unroll
      for (int i = 0; i < NV; ++i) cw += 0.5 * (Ma[i] - fs[i]) * (s.warm[i] - as[i]);
      const bool use_smooth = cw > cs;
#pragma unroll
      for (int i = 0; i < NV; ++i) a[i] = use_smooth ? as[i] : s.warm[i];
    }
    const double scale = 1.0 / (cm.meaninertia * NV);
    double cost = 0;
    for (int iter = 0; iter <= cm.max_iter; ++iter) {
      arrow_mv(M, a, Ma);
      Arrow H = M;
#pragma unroll
      for (int i = 0; i < NV; ++i) fc[i] = 0;
      double newcost = 0;
      for (int r = 0; r < n; ++r) {
        double sj = row_dot(e, r, a) - e.aref[r];
        e.jar[r] = sj;
        if (sj < 0) {
          const double D = e.D[r], f = -D * sj;
          newcost += 0.5 * D * sj * sj;
          const double j0 = e.jr[r][0], j1 = e.jr[r][1], j2 = e.jr[r][2];
          fc[0] += j0 * f; fc[1] += j1 * f; fc[2] += j2 * f;
          H.R[0] += D * j0 * j0; H.R[1] += D * j1 * j0; H.R[2] += D * j1 * j1;
          H.R[3] += D * j2 * j0; H.R[4] += D * j2 * j1; H.R[5] += D * j2 * j2;
          const int l = e.leg[r];
          if (l >= 0) {
            const double l0 = e.jl[r][0], l1 = e.jl[r][1], l2 = e.jl[r][2];
            double* A = l ? H.Af : H.Ab;
            double (*C)[3] = l ? H.Cf : H.Cb;
            double* fl = fc + 3 + 3 * l;
            fl[0] += l0 * f; fl[1] += l1 * f; fl[2] += l2 * f;
            A[0] += D * l0 * l0; A[1] += D * l1 * l0; A[2] += D * l1 * l1;
            A[3] += D * l2 * l0; A[4] += D * l2 * l1; A[5] += D * l2 * l2;
            C[0][0] += D * l0 * j0; C[0][1] += D * l0 * j1; C[0][2] += D * l0 * j2;
            C[1][0] += D * l1 * j0; C[1][1] += D * l1 * j1; C[1][2] += D * l1 * j2;
            C[2][0] += D * l2 * j0; C[2][1] += D * l2 * j1; C[2][2] += D * l2 * j2;
          }
        }
      }
      double g2 = 0;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        newcost += 0.5 * (Ma[i] - fs[i]) * (a[i] - as[i]);
        grad[i] = Ma[i] - fs[i] - fc[i];
        g2 += grad[i] * grad[i];
      }
      const double gnorm = sqrt(g2);
      if (iter > 0) {
        if (scale * (cost - newcost) < cm.tolerance || scale * gnorm < cm.tolerance) break;
      } else if (scale * gnorm < cm.tolerance) {
        break;
      }
      cost = newcost;
      if (iter == cm.max_iter) break;
#pragma unroll
      for (int i = 0; i < NV; ++i) srch[i] = grad[i];
      arrow_solve(H, srch);
#pragma unroll
      for (int i = 0; i < NV; ++i) srch[i] = -srch[i];
      arrow_mv(M, srch, Mv);
      double q1 = 0, q2 = 0;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        q1 += srch[i] * (Ma[i] - fs[i]);
        q2 += srch[i] * Mv[i];
      }
      for (int r = 0; r < n; ++r) e.Jv[r] = row_dot(e, r, srch);
      // stop at |phi'(alpha)| < tolerance * ls_tolerance * |search| / scale (MuJoCo's scaled
      // gradient tolerance of the 1-D problem, ls_tolerance = 0.01).  A tighter test sits
      // below the rounding noise of the row sums and makes single lanes spin to ls_iter
      // while their warp waits (measured: 35 % of all issued instructions at ~2 active lanes).
      double snorm = 0;
#pragma unroll
      for (int i = 0; i < NV; ++i) snorm += srch[i] * srch[i];
      const double gtol = cm.tolerance * 0.01 * sqrt(snorm) / scale;
      double lo = 0, hi = INFINITY, alpha = 0;
      for (int k = 0; k < cm.ls_iter; ++k) {
        double d1 = q1 + alpha * q2, d2 = q2;
        for (int r = 0; r < n; ++r) {
          double x = e.jar[r] + alpha * e.Jv[r];
          if (x < 0) {
            d1 += e.D[r] * x * e.Jv[r];
            d2 += e.D[r] * e.Jv[r] * e.Jv[r];
          }
        }
        if (fabs(d1) < gtol) break;
        if (d1 < 0) lo = alpha; else hi = alpha;
        double next = alpha - d1 / d2;
        if (!(next > lo && next < hi)) next = isinf(hi) ? 2 * alpha + 1 : 0.5 * (lo + hi);
        if (next == alpha) break;
        alpha = next;
      }
      if (alpha == 0) break;
#pragma unroll
      for (int i = 0; i < NV; ++i) a[i] += alpha * srch[i];
    }
  }
  // ---- mj_Euler with implicit joint damping -----------------------------------------------
  {
    Arrow E = M;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      E.Ab[l * (l + 1) / 2 + l] += cm.timestep * cm.damping[3 + l];
      E.Af[l * (l + 1) / 2 + l] += cm.timestep * cm.damping[6 + l];
    }
    double rhs[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) rhs[i] = fs[i] + fc[i];
    arrow_solve(E, rhs);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double vn = s.v[i] + cm.timestep * rhs[i];
      s.v[i] = vn;
      s.q[i] += cm.timestep * vn;
      s.warm[i] = a[i];
    }
  }
}

}  // namespace hct
}  // namespace epb
