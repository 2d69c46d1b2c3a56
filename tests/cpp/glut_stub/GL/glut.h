// tests/cpp/glut_stub/GL/glut.h -- a headless stand-in for <GL/glut.h> (TEST INFRASTRUCTURE).
//
// The reference's demos (self_contact_sparse/main.cpp, internal_cloth_sparse/main.cpp) are GLUT programs; the image
// has no GL.  This header declares the handful of GL / GLUT names those files use as no-ops, so that their UNCHANGED
// sources compile; glutMainLoop (glut_stub.cpp) plays a key script and a number of idle callbacks -- i.e. it drives the
// demo's own StepTime() -- then dumps the demo's aXYZ / aUVW globals and returns.  Nothing is drawn.
//   CLOTHGPU_GLUT_KEYS   key presses delivered before the loop (default "a": start the animation)
//   CLOTHGPU_GLUT_STEPS  idle callbacks to run (default 10)
//   CLOTHGPU_GLUT_DUMP   output file: int32 np, then np*3 doubles aXYZ, then np*3 doubles aUVW
#ifndef CLOTHGPU_GLUT_STUB_H
#define CLOTHGPU_GLUT_STUB_H

typedef float GLfloat;
typedef double GLdouble;
typedef int GLint;
typedef unsigned int GLenum;
typedef unsigned int GLbitfield;
typedef unsigned char GLboolean;

enum {
  GL_COLOR_BUFFER_BIT = 1, GL_DEPTH_BUFFER_BIT = 2, GL_DEPTH_TEST, GL_POLYGON_OFFSET_FILL, GL_MODELVIEW, GL_PROJECTION,
  GL_LIGHTING, GL_LIGHT0, GL_LIGHT1, GL_LIGHT_MODEL_TWO_SIDE, GL_LINES, GL_POINTS, GL_TRIANGLES, GL_FRONT_AND_BACK,
  GL_DIFFUSE, GL_AMBIENT, GL_POSITION, GL_SMOOTH, GL_VIEWPORT,
  GLUT_DOUBLE = 1, GLUT_RGBA = 2, GLUT_DEPTH = 4, GLUT_LEFT_BUTTON = 0, GLUT_ACTIVE_SHIFT = 1, GLUT_ACTIVE_CTRL = 2,
  GLUT_KEY_PAGE_UP = 104, GLUT_KEY_PAGE_DOWN = 105, GLUT_KEY_HOME = 106, GLUT_KEY_END = 107
};

inline void glClearColor(float, float, float, float) {}
inline void glClear(unsigned) {}
inline void glEnable(unsigned) {}
inline void glDisable(unsigned) {}
inline GLboolean glIsEnabled(unsigned) { return 0; }
inline void glPolygonOffset(float, float) {}
inline void glMatrixMode(unsigned) {}
inline void glLoadIdentity() {}
inline void glTranslated(double, double, double) {}
inline void glMultMatrixd(const double*) {}
inline void glLineWidth(float) {}
inline void glPointSize(float) {}
inline void glColor3d(double, double, double) {}
inline void glBegin(unsigned) {}
inline void glEnd() {}
inline void glVertex3d(double, double, double) {}
inline void glVertex3dv(const double*) {}
inline void glNormal3d(double, double, double) {}
inline void glMaterialfv(unsigned, unsigned, const float*) {}
inline void glShadeModel(unsigned) {}
inline void glPushMatrix() {}
inline void glPopMatrix() {}
inline void glLightModelf(unsigned, float) {}
inline void glLightfv(unsigned, unsigned, const float*) {}
inline void glViewport(int, int, int, int) {}
inline void glOrtho(double, double, double, double, double, double) {}
inline void glGetIntegerv(unsigned, GLint* v) { v[0] = 0; v[1] = 0; v[2] = 400; v[3] = 300; }

void glutInit(int* argc, char** argv);
inline void glutInitWindowPosition(int, int) {}
inline void glutInitWindowSize(int, int) {}
inline void glutInitDisplayMode(unsigned) {}
inline int glutCreateWindow(const char*) { return 1; }
inline void glutPostRedisplay() {}
inline void glutSwapBuffers() {}
inline void glutWireSphere(double, int, int) {}
inline int glutGetModifiers() { return 0; }
void glutDisplayFunc(void (*f)(void));
void glutIdleFunc(void (*f)(void));
void glutReshapeFunc(void (*f)(int, int));
void glutMotionFunc(void (*f)(int, int));
void glutMouseFunc(void (*f)(int, int, int, int));
void glutKeyboardFunc(void (*f)(unsigned char, int, int));
void glutSpecialFunc(void (*f)(int, int, int));
void glutMainLoop();

#endif
