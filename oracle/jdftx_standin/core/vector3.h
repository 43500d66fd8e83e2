// Stand-in for JDFTx core/vector3.h (JDFTx is an external, un-vendored dependency of the
// reference: CMakeLists.txt:4-5,43 take user-supplied JDFTX_SRC/JDFTX_BUILD paths; no pinned version).
// TEST INFRASTRUCTURE ONLY: exists so that the UNMODIFIED reference sources under /root/reference
// compile into oracle/_ref. Written from the published interface, not copied.
//
// Arithmetic contract (every operation is a single IEEE-754 double op, evaluated left to right):
//   length()      = sqrt(v0*v0 + v1*v1 + v2*v2)
//   normalize(v)  = v * (1.0 / v.length())
//   s*v, v*s      = component-wise product
//   cross(a,b)    = (a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0)
#ifndef STANDIN_CORE_VECTOR3_H
#define STANDIN_CORE_VECTOR3_H

#include <cmath>
#include <vector>
#include <cstdio>
#include <cstdlib>

template<typename T = double> class vector3
{
	T c_[3];
public:
	explicit vector3(T a = 0, T b = 0, T c = 0) { c_[0] = a; c_[1] = b; c_[2] = c; }
	T& operator[](int k) { return c_[k]; }
	const T& operator[](int k) const { return c_[k]; }
	T& x() { return c_[0]; }
	T& y() { return c_[1]; }
	T& z() { return c_[2]; }
	const T& x() const { return c_[0]; }
	const T& y() const { return c_[1]; }
	const T& z() const { return c_[2]; }

	vector3 operator+(const vector3& o) const { return vector3(c_[0] + o.c_[0], c_[1] + o.c_[1], c_[2] + o.c_[2]); }
	vector3 operator-(const vector3& o) const { return vector3(c_[0] - o.c_[0], c_[1] - o.c_[1], c_[2] - o.c_[2]); }
	vector3 operator-() const { return vector3(-c_[0], -c_[1], -c_[2]); }
	vector3& operator+=(const vector3& o) { c_[0] += o.c_[0]; c_[1] += o.c_[1]; c_[2] += o.c_[2]; return *this; }
	vector3& operator-=(const vector3& o) { c_[0] -= o.c_[0]; c_[1] -= o.c_[1]; c_[2] -= o.c_[2]; return *this; }
	vector3 operator*(T s) const { return vector3(c_[0] * s, c_[1] * s, c_[2] * s); }
	vector3& operator*=(T s) { c_[0] *= s; c_[1] *= s; c_[2] *= s; return *this; }

	T length_squared() const { return c_[0] * c_[0] + c_[1] * c_[1] + c_[2] * c_[2]; }
	T length() const { return sqrt(length_squared()); }
};

template<typename T> vector3<T> operator*(T s, const vector3<T>& v) { return vector3<T>(v[0] * s, v[1] * s, v[2] * s); }
// mixed int*double etc. are not used by the reference; double literal * vector3<> resolves to the template above.

template<typename T> T dot(const vector3<T>& a, const vector3<T>& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

template<typename T> vector3<T> cross(const vector3<T>& a, const vector3<T>& b)
{	return vector3<T>(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}

inline vector3<> normalize(const vector3<>& v) { return v * (1.0 / v.length()); }

#endif
