// tests/cpp/glut_stub/glut_stub.cpp -- the event loop of the headless GLUT stand-in (see GL/glut.h here): delivers a
// key script, runs N idle callbacks (each one is a StepTime() of the demo while its animation flag is on), draws one
// frame through the demo's display callback (all GL calls are no-ops) and dumps the demo's state vectors.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "GL/glut.h"

// the demo's own globals (self_contact_sparse/main.cpp:141-142, internal_cloth_sparse/main.cpp:141-142)
extern std::vector<double> aXYZ;
extern std::vector<double> aUVW;

namespace {
void (*g_display)(void) = nullptr;
void (*g_idle)(void) = nullptr;
void (*g_keyboard)(unsigned char, int, int) = nullptr;
}  // namespace

void glutInit(int*, char**) {}
void glutDisplayFunc(void (*f)(void)) { g_display = f; }
void glutIdleFunc(void (*f)(void)) { g_idle = f; }
void glutReshapeFunc(void (*)(int, int)) {}
void glutMotionFunc(void (*)(int, int)) {}
void glutMouseFunc(void (*)(int, int, int, int)) {}
void glutKeyboardFunc(void (*f)(unsigned char, int, int)) { g_keyboard = f; }
void glutSpecialFunc(void (*)(int, int, int)) {}

void glutMainLoop() {
  const char* keys = getenv("CLOTHGPU_GLUT_KEYS");
  if (!keys) keys = "a";
  const char* st = getenv("CLOTHGPU_GLUT_STEPS");
  const int steps = st ? atoi(st) : 10;
  if (g_keyboard)
    for (const char* k = keys; *k; k++) g_keyboard((unsigned char)*k, 0, 0);
  for (int i = 0; i < steps && g_idle; i++) g_idle();
  if (g_display) g_display();
  const char* path = getenv("CLOTHGPU_GLUT_DUMP");
  if (path) {
    FILE* f = fopen(path, "wb");
    if (!f) { perror(path); exit(2); }
    const int np = (int)aXYZ.size() / 3;
    fwrite(&np, sizeof(int), 1, f);
    fwrite(aXYZ.data(), sizeof(double), aXYZ.size(), f);
    fwrite(aUVW.data(), sizeof(double), aUVW.size(), f);
    fclose(f);
  }
  printf("glut_stub: %d steps, %d vertices\n", steps, (int)aXYZ.size() / 3);
}
