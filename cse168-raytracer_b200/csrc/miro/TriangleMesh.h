// TriangleMesh.h -- indexed triangle mesh + OBJ ingest of the host API layer (reference TriangleMesh.h,
// TriangleMeshLoad.cpp).  load() keeps the reference loader's observable semantics: 80-byte
// line reads, first three face tokens, v/vt/vn index forms, ctm on positions and the normalised inverse
// transpose on normals, per-face normals synthesised when the third token carries none and then averaged
// per vertex (starting from the default-constructed Vector3).
#ifndef MIROHOST_TRIANGLEMESH_H
#define MIROHOST_TRIANGLEMESH_H
#include <cstdio>
#include "Matrix4x4.h"

class TriangleMesh {
public:
    TriangleMesh();
    ~TriangleMesh();
    bool load(const char* file, const Matrix4x4& ctm = Matrix4x4());
    void createSingleTriangle();
    void setV1(const Vector3& v) { m_vertices[0] = v; }
    void setV2(const Vector3& v) { m_vertices[1] = v; }
    void setV3(const Vector3& v) { m_vertices[2] = v; }
    void setN1(const Vector3& n) { m_normals[0] = n; }
    void setN2(const Vector3& n) { m_normals[1] = n; }
    void setN3(const Vector3& n) { m_normals[2] = n; }
    struct TupleI3 { unsigned int v[3]; };
    struct VectorR2 { float x, y; };
    Vector3* vertices() { return m_vertices; }
    Vector3* normals() { return m_normals; }
    TupleI3* vIndices() { return m_vertexIndices; }
    TupleI3* nIndices() { return m_normalIndices; }
    VectorR2* texCoords() { return m_texCoords; }
    TupleI3* tIndices() { return m_texCoordIndices; }
    int numTris() { return m_numTris; }
    int numTextCoords() { return m_numTextCoords; }
protected:
    void loadObj(FILE* fp, const Matrix4x4& ctm);
    Vector3* m_normals;
    Vector3* m_vertices;
    TupleI3* m_normalIndices;
    TupleI3* m_vertexIndices;
    VectorR2* m_texCoords;
    TupleI3* m_texCoordIndices;
    unsigned int m_numVertices, m_numTris, m_numTextCoords;
};
#endif
