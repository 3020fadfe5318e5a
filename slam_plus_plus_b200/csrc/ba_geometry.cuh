// ba_geometry.cuh -- SE(3) / pinhole device math of the BA and SE(3) pose-graph kernels.
//
// Operation order follows the reference so that the forward-difference Jacobians (delta = 1e-9) land
// inside the reference's own build-to-build envelope:
//   axis-angle <-> quaternion   include/slam/3DSolverBase.h:476-519 (f_AxisAngle_to_Quat), :556-649 (f_Quat_to_AxisAngle)
//   pose composition            include/slam/3DSolverBase.h:806-849 (Relative_to_Absolute), :892-946 (Absolute_to_Relative)
//   projection                  include/slam/BASolverBase.h:260-327 (Project_P2C)
// Quaternion algebra is written out the way Eigen evaluates it (Quaternion product, _transformVector,
// toRotationMatrix), since those are what the reference calls.
#pragma once

#include <math.h>

namespace spp {

struct Quat { double w, x, y, z; };

__device__ __forceinline__ double norm3(double a, double b, double c)
{
	return sqrt(a * a + b * b + c * c);
}

__device__ __forceinline__ void quat_normalize(Quat &q)
{
	double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
	q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}

// 3DSolverBase.h:476-519
__device__ __forceinline__ void axis_angle_to_quat(double ax, double ay, double az, Quat &q)
{
	double f_angle = norm3(ax, ay, az);
	if(f_angle < 1e-12) {
		q.w = cos(f_angle * .5);
		q.x = ax * .5; q.y = ay * .5; q.z = az * .5;
		quat_normalize(q);
	} else {
		double f_half_angle = f_angle * .5;
		double s, c;
		sincos(f_half_angle, &s, &c);
		double f_q = s / f_angle;
		if(c < 0) {
			c = -c;
			f_q = -f_q;
		}
		q.w = c;
		q.x = ax * f_q; q.y = ay * f_q; q.z = az * f_q;
		if(c > 1 - 1e-6)
			quat_normalize(q);
	}
}

// 3DSolverBase.h:556-649 (the "norm and atan and atan2" variant that is compiled in)
__device__ __forceinline__ void quat_to_axis_angle(const Quat &q, double &ax, double &ay, double &az)
{
	const double f_w = q.w;
	const double f_abs_w = fabs(f_w), f_norm = norm3(q.x, q.y, q.z);
	const double f_abs_half = (f_abs_w > 1e-3)? atan(f_norm / f_abs_w) : atan2(f_norm, f_abs_w);
	const double f_half = copysign(f_abs_half, f_w);
	if(f_norm < 1e-12) {
		ax = q.x * 2.0; ay = q.y * 2.0; az = q.z * 2.0;
	} else {
		double f_s = f_half * 2 / f_norm;
		ax = q.x * f_s; ay = q.y * f_s; az = q.z * f_s;
	}
}

// Eigen::Quaternion product a * b
__device__ __forceinline__ Quat quat_mul(const Quat &a, const Quat &b)
{
	Quat r;
	r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
	r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
	r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
	r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
	return r;
}

__device__ __forceinline__ Quat quat_conj(const Quat &a)
{
	Quat r;
	r.w = a.w; r.x = -a.x; r.y = -a.y; r.z = -a.z;
	return r;
}

// Eigen::QuaternionBase::_transformVector: v + w * uv + vec x uv with uv = 2 (vec x v)
__device__ __forceinline__ void quat_rotate(const Quat &q, double vx, double vy, double vz,
	double &rx, double &ry, double &rz)
{
	double ux = q.y * vz - q.z * vy, uy = q.z * vx - q.x * vz, uz = q.x * vy - q.y * vx;
	ux += ux; uy += uy; uz += uz;
	rx = vx + q.w * ux + (q.y * uz - q.z * uy);
	ry = vy + q.w * uy + (q.z * ux - q.x * uz);
	rz = vz + q.w * uz + (q.x * uy - q.y * ux);
}

// Eigen::QuaternionBase::toRotationMatrix, row-major R[9]
__device__ __forceinline__ void quat_to_rotmat(const Quat &q, double *R)
{
	const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
	const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
	const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
	const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
	R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
	R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
	R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

// 3DSolverBase.h:806-849: dest = v1 (+) v2, poses are [t(3), axis-angle(3)]
__device__ __forceinline__ void relative_to_absolute(const double *v1, const double *v2, double *dest)
{
	Quat q1, q2;
	axis_angle_to_quat(v1[3], v1[4], v1[5], q1);
	axis_angle_to_quat(v2[3], v2[4], v2[5], q2);
	double rx, ry, rz;
	quat_rotate(q1, v2[0], v2[1], v2[2], rx, ry, rz);
	double d0 = v1[0] + rx, d1 = v1[1] + ry, d2 = v1[2] + rz;
	Quat q = quat_mul(q1, q2);
	double ax, ay, az;
	quat_to_axis_angle(q, ax, ay, az);
	dest[0] = d0; dest[1] = d1; dest[2] = d2;
	dest[3] = ax; dest[4] = ay; dest[5] = az;
}

// 3DSolverBase.h:892-946: dest = v1^-1 (+) v2
__device__ __forceinline__ void absolute_to_relative(const double *v1, const double *v2, double *dest)
{
	Quat q1, q2;
	axis_angle_to_quat(v1[3], v1[4], v1[5], q1);
	axis_angle_to_quat(v2[3], v2[4], v2[5], q2);
	Quat q1i = quat_conj(q1);
	double rx, ry, rz;
	quat_rotate(q1i, v2[0] - v1[0], v2[1] - v1[1], v2[2] - v1[2], rx, ry, rz);
	Quat q = quat_mul(q1i, q2);
	double ax, ay, az;
	quat_to_axis_angle(q, ax, ay, az);
	dest[0] = rx; dest[1] = ry; dest[2] = rz;
	dest[3] = ax; dest[4] = ay; dest[5] = az;
}

// [R | t] (12 doubles: R row-major, then t) of a camera pose [t, axis-angle]
__device__ __forceinline__ void pose_to_Rt(const double *pose, double *Rt)
{
	Quat q;
	axis_angle_to_quat(pose[3], pose[4], pose[5], q);
	quat_to_rotmat(q, Rt);
	Rt[9] = pose[0]; Rt[10] = pose[1]; Rt[11] = pose[2];
}

// BASolverBase.h:260-327 with [R|t] already formed; intr = fx, fy, cx, cy, k (k = d / (.5 (fx + fy)))
__device__ __forceinline__ void project_Rt(const double *Rt, double fx, double fy, double cx, double cy, double k,
	double X, double Y, double Z, double &u, double &v)
{
	// x = Rt * [X; 1]
	double x0 = Rt[0] * X + Rt[1] * Y + Rt[2] * Z + Rt[9];
	double x1 = Rt[3] * X + Rt[4] * Y + Rt[5] * Z + Rt[10];
	double x2 = Rt[6] * X + Rt[7] * Y + Rt[8] * Z + Rt[11];
	// uv = A * x; uv /= uv(2)
	double u0 = fx * x0 + cx * x2, u1 = fy * x1 + cy * x2;
	u0 /= x2; u1 /= x2;
	// radial distortion around the principal point
	double dx = u0 - cx, dy = u1 - cy;
	double r = sqrt(dx * dx + dy * dy);
	double f_s = 1 + r * r * k;
	u = cx + f_s * dx;
	v = cy + f_s * dy;
}

} // namespace spp
