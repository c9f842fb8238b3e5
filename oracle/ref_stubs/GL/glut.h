/* Stand-in for <GL/glut.h>, used ONLY to compile the unmodified reference
 * sources under /root/reference into oracle/_ref (test infrastructure).
 * The GPU box and this container have no OpenGL/GLUT headers; the hot path
 * (BVH / Triangle / Scene::trace) never calls any of these, so every entry
 * point is an empty variadic inline.  Not product code. */
#ifndef MIRO_ORACLE_GLUT_STUB_H
#define MIRO_ORACLE_GLUT_STUB_H
typedef unsigned int GLenum;
typedef float GLfloat;
typedef int GLint;
typedef int GLsizei;
typedef double GLdouble;
typedef unsigned int GLbitfield;
typedef unsigned char GLubyte;
enum {
    GL_COLOR_BUFFER_BIT = 1, GL_DEPTH_BUFFER_BIT = 2, GL_TRIANGLES = 4, GL_QUADS = 7,
    GL_BACK = 10, GL_FRONT = 11, GL_PROJECTION = 12, GL_MODELVIEW = 13, GL_RGB = 14,
    GL_UNSIGNED_BYTE = 15, GL_LIGHTING = 16, GL_TEXTURE_2D = 17, GL_SMOOTH = 18,
    GL_FRONT_AND_BACK = 19, GL_LINE = 20, GL_FLAT = 21, GL_DEPTH_TEST = 22, GL_FILL = 23,
    GLUT_RGB = 0, GLUT_DOUBLE = 2, GLUT_DEPTH = 16, GLUT_LEFT_BUTTON = 0, GLUT_MIDDLE_BUTTON = 1,
    GLUT_RIGHT_BUTTON = 2, GLUT_DOWN = 0, GLUT_UP = 1
};
#define MIRO_GL_NOOP(name) static inline void name(...) {}
MIRO_GL_NOOP(glClear) MIRO_GL_NOOP(glutSwapBuffers) MIRO_GL_NOOP(glBegin) MIRO_GL_NOOP(glEnd)
MIRO_GL_NOOP(glVertex3f) MIRO_GL_NOOP(glColor3f) MIRO_GL_NOOP(glPushMatrix) MIRO_GL_NOOP(glPopMatrix)
MIRO_GL_NOOP(glTranslatef) MIRO_GL_NOOP(glutWireSphere) MIRO_GL_NOOP(glDrawBuffer)
MIRO_GL_NOOP(glMatrixMode) MIRO_GL_NOOP(glLoadIdentity) MIRO_GL_NOOP(gluPerspective)
MIRO_GL_NOOP(gluLookAt) MIRO_GL_NOOP(glRasterPos2f) MIRO_GL_NOOP(glDrawPixels) MIRO_GL_NOOP(glFinish)
MIRO_GL_NOOP(glutPostRedisplay) MIRO_GL_NOOP(glutInit) MIRO_GL_NOOP(glutInitWindowSize)
MIRO_GL_NOOP(glutInitDisplayMode) MIRO_GL_NOOP(glutInitWindowPosition) MIRO_GL_NOOP(glutCreateWindow)
MIRO_GL_NOOP(glClearColor) MIRO_GL_NOOP(glDisable) MIRO_GL_NOOP(glEnable) MIRO_GL_NOOP(glShadeModel)
MIRO_GL_NOOP(glPolygonMode) MIRO_GL_NOOP(glutDisplayFunc) MIRO_GL_NOOP(glutKeyboardFunc)
MIRO_GL_NOOP(glutMouseFunc) MIRO_GL_NOOP(glutMotionFunc) MIRO_GL_NOOP(glutReshapeFunc)
MIRO_GL_NOOP(glutMainLoop) MIRO_GL_NOOP(glReadPixels) MIRO_GL_NOOP(glViewport)
#undef MIRO_GL_NOOP
#endif
