// Vector3.h -- binary32 3-vector of the host API layer.  Interface and rounding behaviour follow the
// reference's Vector3.h (division multiplies by a rounded reciprocal, :125-129; dot sums left to right,
// :243-246; the default constructor yields (0,1,2), :26-27 -- the OBJ loader's normal averaging depends on it).
#ifndef MIROHOST_VECTOR3_H
#define MIROHOST_VECTOR3_H
#include <cmath>
#include <iostream>

struct VectorR2 { float x, y; };

class Vector3 {
public:
    float x, y, z;
    Vector3() : x(0), y(1), z(2) {}
    Vector3(float s) : x(s), y(s), z(s) {}
    Vector3(float a, float b, float c) : x(a), y(b), z(c) {}
    void set(float a) { x = y = z = a; }
    void set(float a, float b, float c) { x = a; y = b; z = c; }
    void set(const Vector3& v) { x = v.x; y = v.y; z = v.z; }
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
    Vector3 operator+(const Vector3& v) const { return Vector3(x + v.x, y + v.y, z + v.z); }
    Vector3 operator-(const Vector3& v) const { return Vector3(x - v.x, y - v.y, z - v.z); }
    Vector3 operator-() const { return Vector3(-x, -y, -z); }
    Vector3 operator*(float a) const { return Vector3(x * a, y * a, z * a); }
    Vector3 operator*(const Vector3& v) const { return Vector3(x * v.x, y * v.y, z * v.z); }
    Vector3 operator/(float a) const { const float r = float(1) / a; return Vector3(x * r, y * r, z * r); }
    const Vector3& operator+=(const Vector3& v) { x += v.x; y += v.y; z += v.z; return *this; }
    const Vector3& operator-=(const Vector3& v) { x -= v.x; y -= v.y; z -= v.z; return *this; }
    const Vector3& operator*=(float a) { x *= a; y *= a; z *= a; return *this; }
    const Vector3& operator/=(float a) { const float r = float(1) / a; x *= r; y *= r; z *= r; return *this; }
    bool operator==(const Vector3& v) const { return v.x == x && v.y == y && v.z == z; }
    bool operator!=(const Vector3& v) const { return !(*this == v); }
    float length2() const { return x * x + y * y + z * z; }
    float length() const { return sqrtf(length2()); }
    const Vector3& normalize() { return (*this /= length()); }
    Vector3 normalized() const { return *this / length(); }
    float average() const { return (x + y + z) / 3.0f; }
};

inline Vector3 operator*(float s, const Vector3& v) { return Vector3(v.x * s, v.y * s, v.z * s); }
inline float dot(const Vector3& a, const Vector3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vector3 cross(const Vector3& a, const Vector3& b)
{
    return Vector3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline std::ostream& operator<<(std::ostream& o, const Vector3& v) { return o << v.x << ",\t" << v.y << ",\t" << v.z; }
#endif
