// TEST INFRASTRUCTURE ONLY (see orc_real.hpp).
// Small vector / quaternion / rotation algebra, restating
//   lib/FlightPhysics/src/quaternions.jl:109-115 (Hamilton product)
//   lib/FlightPhysics/src/attitude.jl:98-103 (rotate), :118 (dt), :304-308
//   (axis-angle -> quat), :382-395 (quat <-> Euler), :175-233 (RMatrix <-> quat),
//   :436-474 (Euler rates), :478 (wrap_to_pi)
// in the reference's operation order.
#pragma once
#include "orc_real.hpp"

static const double ORC_PI = 3.14159265358979323846;

struct V3 {
    real x, y, z;
    V3() : x(0.0), y(0.0), z(0.0) {}
    V3(real a, real b, real c) : x(a), y(b), z(c) {}
    real operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(real s, V3 a) { return V3(s * a.x, s * a.y, s * a.z); }
inline V3 operator*(V3 a, real s) { return s * a; }
inline V3 operator/(V3 a, real s) { return V3(a.x / s, a.y / s, a.z / s); }
inline real dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline real norm(V3 a) { return sqrt(dot(a, a)); }
inline V3 normalize(V3 a) { return a / norm(a); }

struct M3 {  // row-major 3x3
    real m[3][3];
    M3() { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = 0.0; }
};
inline M3 operator+(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }
inline M3 operator-(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] - b.m[i][j]; return r; }
inline M3 operator*(real s, const M3& a) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = s * a.m[i][j]; return r; }
inline M3 operator*(const M3& a, const M3& b) {
    M3 r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        real s = a.m[i][0] * b.m[0][j];
        s = s + a.m[i][1] * b.m[1][j];
        s = s + a.m[i][2] * b.m[2][j];
        r.m[i][j] = s;
    }
    return r;
}
inline V3 operator*(const M3& a, V3 v) {
    return V3(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
              a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
              a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z);
}
inline M3 transpose(const M3& a) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i]; return r; }
// attitude.jl:36-41
inline M3 v2skew(V3 v) {
    M3 r;
    r.m[0][1] = -v.z; r.m[0][2] = v.y;
    r.m[1][0] = v.z;  r.m[1][2] = -v.x;
    r.m[2][0] = -v.y; r.m[2][1] = v.x;
    return r;
}
// StaticArrays 3x3 `\` : solved here by cofactor expansion (agrees with LU to rounding).
inline V3 solve3(const M3& A, V3 b) {
    const real (*a)[3] = A.m;
    real c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1];
    real c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2];
    real c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
    real det = a[0][0] * c00 + a[0][1] * c01 + a[0][2] * c02;
    real c10 = a[0][2] * a[2][1] - a[0][1] * a[2][2];
    real c11 = a[0][0] * a[2][2] - a[0][2] * a[2][0];
    real c12 = a[0][1] * a[2][0] - a[0][0] * a[2][1];
    real c20 = a[0][1] * a[1][2] - a[0][2] * a[1][1];
    real c21 = a[0][2] * a[1][0] - a[0][0] * a[1][2];
    real c22 = a[0][0] * a[1][1] - a[0][1] * a[1][0];
    real id = 1.0 / det;
    return V3((c00 * b.x + c10 * b.y + c20 * b.z) * id,
              (c01 * b.x + c11 * b.y + c21 * b.z) * id,
              (c02 * b.x + c12 * b.y + c22 * b.z) * id);
}

struct Q4 {  // real part w, imaginary (x,y,z)
    real w, x, y, z;
    Q4() : w(1.0), x(0.0), y(0.0), z(0.0) {}
    Q4(real a, real b, real c, real d) : w(a), x(b), y(c), z(d) {}
    V3 im() const { return V3(x, y, z); }
};
// quaternions.jl:109-115
inline Q4 qmul(Q4 a, Q4 b) {
    real re = a.w * b.w - dot(a.im(), b.im());
    V3 im = a.w * b.im() + b.w * a.im() + cross(a.im(), b.im());
    return Q4(re, im.x, im.y, im.z);
}
inline Q4 qconj(Q4 a) { return Q4(a.w, -a.x, -a.y, -a.z); }
inline real qnorm(Q4 a) { return sqrt(a.w * a.w + a.x * a.x + a.y * a.y + a.z * a.z); }
inline Q4 qnormalize(Q4 a) { real n = qnorm(a); return Q4(a.w / n, a.x / n, a.y / n, a.z / n); }
// attitude.jl:98-103
inline V3 qrot(Q4 q, V3 v) {
    V3 qi = q.im();
    return v + cross(2.0 * qi, q.w * v + cross(qi, v));
}
inline V3 qrot_inv(Q4 q, V3 v) { return qrot(qconj(q), v); }
// attitude.jl:118 : 0.5 * (q * FreeQuat(imag = w))
inline Q4 qdt(Q4 q, V3 w) {
    Q4 p = qmul(q, Q4(0.0, w.x, w.y, w.z));
    return Q4(0.5 * p.w, 0.5 * p.x, 0.5 * p.y, 0.5 * p.z);
}
// attitude.jl:248-250, 304-308
inline Q4 Rz(real a) { real c = cos(0.5 * a), s = sin(0.5 * a); return Q4(c, 0.0 * s, 0.0 * s, 1.0 * s); }
inline Q4 Ry(real a) { real c = cos(0.5 * a), s = sin(0.5 * a); return Q4(c, 0.0 * s, 1.0 * s, 0.0 * s); }
inline Q4 Rx(real a) { real c = cos(0.5 * a), s = sin(0.5 * a); return Q4(c, 1.0 * s, 0.0 * s, 0.0 * s); }

struct Euler { real psi, theta, phi; Euler() : psi(0.0), theta(0.0), phi(0.0) {} Euler(real a, real b, real c) : psi(a), theta(b), phi(c) {} };
// attitude.jl:382-390
inline Euler euler_from_q(Q4 q) {
    real q1 = q.w, q2 = q.x, q3 = q.y, q4 = q.z;
    real s2 = q2 * q2, s3 = q3 * q3, s4 = q4 * q4;
    real psi = atan2(2.0 * (q1 * q4 + q2 * q3), 1.0 - 2.0 * (s3 + s4));
    real theta = asin(rclamp(2.0 * (q1 * q3 - q2 * q4), -1.0, 1.0));
    real phi = atan2(2.0 * (q1 * q2 + q3 * q4), 1.0 - 2.0 * (s2 + s3));
    return Euler(psi, theta, phi);
}
// attitude.jl:392-395
inline Q4 q_from_euler(Euler e) { return qmul(qmul(Rz(e.psi), Ry(e.theta)), Rx(e.phi)); }
// attitude.jl:175-187 (RQuat -> RMatrix)
inline M3 rmat_from_q(Q4 r) {
    Q4 q = qnormalize(r);
    real s2 = q.x * q.x, s3 = q.y * q.y, s4 = q.z * q.z;
    real dq12 = 2.0 * q.w * q.x, dq13 = 2.0 * q.w * q.y, dq14 = 2.0 * q.w * q.z;
    real dq23 = 2.0 * q.x * q.y, dq24 = 2.0 * q.x * q.z, dq34 = 2.0 * q.y * q.z;
    M3 M;
    M.m[0][0] = 1.0 - 2.0 * (s3 + s4); M.m[0][1] = dq23 - dq14; M.m[0][2] = dq24 + dq13;
    M.m[1][0] = dq23 + dq14; M.m[1][1] = 1.0 - 2.0 * (s2 + s4); M.m[1][2] = dq34 - dq12;
    M.m[2][0] = dq24 - dq13; M.m[2][1] = dq34 + dq12; M.m[2][2] = 1.0 - 2.0 * (s2 + s3);
    return M;
}
// attitude.jl:189-233 (RMatrix -> RQuat)
inline Q4 q_from_rmat(const M3& R) {
    real tr = R.m[0][0] + R.m[1][1] + R.m[2][2];
    real cand[4] = {tr, R.m[0][0], R.m[1][1], R.m[2][2]};
    int imax = 0;
    for (int i = 1; i < 4; i++) if (cand[i] > cand[imax]) imax = i;
    Q4 v;
    if (imax == 0) v = Q4(1.0 + tr, R.m[2][1] - R.m[1][2], R.m[0][2] - R.m[2][0], R.m[1][0] - R.m[0][1]);
    else if (imax == 1) v = Q4(R.m[2][1] - R.m[1][2], 1.0 + 2.0 * R.m[0][0] - tr, R.m[0][1] + R.m[1][0], R.m[2][0] + R.m[0][2]);
    else if (imax == 2) v = Q4(R.m[0][2] - R.m[2][0], R.m[0][1] + R.m[1][0], 1.0 + 2.0 * R.m[1][1] - tr, R.m[1][2] + R.m[2][1]);
    else v = Q4(R.m[1][0] - R.m[0][1], R.m[2][0] + R.m[0][2], R.m[1][2] + R.m[2][1], 1.0 + 2.0 * R.m[2][2] - tr);
    return qnormalize(v);
}
// attitude.jl:436-449 (Euler angle rates)
inline V3 euler_dt(Euler e, V3 w) {
    real sphi = sin(e.phi), cphi = cos(e.phi);
    real tth = tan(e.theta), sec = 1.0 / cos(e.theta);
    return V3(sphi * sec * w.y + cphi * sec * w.z,
              cphi * w.y - sphi * w.z,
              w.x + sphi * tth * w.y + cphi * tth * w.z);
}
// attitude.jl:460-474 (angular velocity from Euler rates)
inline V3 omega_from_euler_dt(Euler e, V3 ed) {
    real sth = sin(e.theta), cth = cos(e.theta), sphi = sin(e.phi), cphi = cos(e.phi);
    return V3(-sth * ed.x + ed.z,
              cth * sphi * ed.x + cphi * ed.y,
              cth * cphi * ed.x - sphi * ed.y);
}
inline real azimuth(V3 v) { return atan2(v.y, v.x); }                                   // attitude.jl:476
inline real inclination(V3 v) { return atan2(-v.z, sqrt(v.x * v.x + v.y * v.y)); }     // attitude.jl:477
inline real wrap_to_pi(real x) { return x + 2.0 * ORC_PI * floor((ORC_PI - x) / (2.0 * ORC_PI)); }  // :478
