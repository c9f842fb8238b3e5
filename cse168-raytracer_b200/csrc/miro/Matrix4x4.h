// Matrix4x4.h -- 4x4 transform of the host API layer (m_ij = row i, column j; reference Matrix4x4.h).
// Only what the geometry ingest needs: products, the affine point transform that ignores row 4 (:581-588),
// transpose, and the cofactor inverse whose reciprocal determinant is formed in double (:349).
#ifndef MIROHOST_MATRIX4X4_H
#define MIROHOST_MATRIX4X4_H
#include "Vector3.h"

struct Vector4 {
    float x, y, z, w;
    Vector4() : x(0), y(0), z(0), w(0) {}
    Vector4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};

class Matrix4x4 {
public:
    float m11, m12, m13, m14, m21, m22, m23, m24, m31, m32, m33, m34, m41, m42, m43, m44;
    Matrix4x4() { setIdentity(); }
    // sixteen values in ROW order
    Matrix4x4(float a11, float a12, float a13, float a14, float a21, float a22, float a23, float a24,
              float a31, float a32, float a33, float a34, float a41, float a42, float a43, float a44)
    {
        set(a11, a12, a13, a14, a21, a22, a23, a24, a31, a32, a33, a34, a41, a42, a43, a44);
    }
    // four COLUMN vectors
    Matrix4x4(const Vector4& c1, const Vector4& c2, const Vector4& c3, const Vector4& c4)
    {
        set(c1.x, c2.x, c3.x, c4.x, c1.y, c2.y, c3.y, c4.y, c1.z, c2.z, c3.z, c4.z, c1.w, c2.w, c3.w, c4.w);
    }
    void set(float a11, float a12, float a13, float a14, float a21, float a22, float a23, float a24,
             float a31, float a32, float a33, float a34, float a41, float a42, float a43, float a44)
    {
        m11 = a11; m12 = a12; m13 = a13; m14 = a14; m21 = a21; m22 = a22; m23 = a23; m24 = a24;
        m31 = a31; m32 = a32; m33 = a33; m34 = a34; m41 = a41; m42 = a42; m43 = a43; m44 = a44;
    }
    void setIdentity() { set(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1); }
    void setColumn4(const Vector4& c) { m14 = c.x; m24 = c.y; m34 = c.z; m44 = c.w; }
    float* rows() { return &m11; }
    const float* rows() const { return &m11; }

    void transpose()
    {
        std::swap(m12, m21); std::swap(m13, m31); std::swap(m14, m41);
        std::swap(m23, m32); std::swap(m24, m42); std::swap(m34, m43);
    }

    Matrix4x4& invert()
    {
        const float* a = rows();
        // 2x2 minors of row pairs (3,4), (2,4), (2,3): index [pair][column pair]
        auto A = [a](int r, int c) { return a[4 * r + c]; };
        auto minor2 = [&](int r0, int r1, int c0, int c1) { return A(r0, c0) * A(r1, c1) - A(r0, c1) * A(r1, c0); };
        const int cp[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
        float t34[6], t24[6], t23[6];
        for (int k = 0; k < 6; ++k) {
            t34[k] = minor2(2, 3, cp[k][0], cp[k][1]);
            t24[k] = minor2(1, 3, cp[k][0], cp[k][1]);
            t23[k] = minor2(1, 2, cp[k][0], cp[k][1]);
        }
        // 3x3 minor of the matrix with row r removed (rows r1 + the pair), expanded along row r1 over columns != c
        auto minor3 = [&](int r1, const float* pair, int c) {
            int cols[3], n = 0;
            for (int j = 0; j < 4; ++j) if (j != c) cols[n++] = j;
            auto idx = [&](int ca, int cb) { for (int k = 0; k < 6; ++k) if (cp[k][0] == ca && cp[k][1] == cb) return k; return 0; };
            return A(r1, cols[0]) * pair[idx(cols[1], cols[2])] - A(r1, cols[1]) * pair[idx(cols[0], cols[2])] +
                   A(r1, cols[2]) * pair[idx(cols[0], cols[1])];
        };
        float sd[4][4];  // sd[i][j]: minor with row i, column j removed
        for (int j = 0; j < 4; ++j) {
            sd[0][j] = minor3(1, t34, j);
            sd[1][j] = minor3(0, t34, j);
            sd[2][j] = minor3(0, t24, j);
            sd[3][j] = minor3(0, t23, j);
        }
        const float detInv = 1.0 / (m11 * sd[0][0] - m12 * sd[0][1] + m13 * sd[0][2] - m14 * sd[0][3]);
        float r[16];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                const float s = sd[j][i];
                r[4 * i + j] = ((i + j) & 1) ? -s * detInv : s * detInv;
            }
        for (int k = 0; k < 16; ++k) rows()[k] = r[k];
        return *this;
    }

    Matrix4x4& operator*=(const Matrix4x4& B)
    {
        const float* a = rows(); const float* b = B.rows();
        float r[16];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                r[4 * i + j] = a[4 * i] * b[j] + a[4 * i + 1] * b[4 + j] + a[4 * i + 2] * b[8 + j] + a[4 * i + 3] * b[12 + j];
        for (int k = 0; k < 16; ++k) rows()[k] = r[k];
        return *this;
    }
};

inline Matrix4x4 operator*(const Matrix4x4& A, const Matrix4x4& B) { Matrix4x4 r = A; r *= B; return r; }
// Point transform; the fourth row is ignored.
inline Vector3 operator*(const Matrix4x4& A, const Vector3& u)
{
    return Vector3(A.m11 * u.x + A.m12 * u.y + A.m13 * u.z + A.m14, A.m21 * u.x + A.m22 * u.y + A.m23 * u.z + A.m24,
                   A.m31 * u.x + A.m32 * u.y + A.m33 * u.z + A.m34);
}
#endif
