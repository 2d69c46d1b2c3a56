// cli/clothsim.cpp -- headless command-line front end of the two sparse demos of the reference
// (self_contact_sparse/main.cpp, internal_cloth_sparse/main.cpp) over the C ABI (include/clothgpu.h).
//
// What the GLUT demos do interactively, this does from the command line and reproducibly: build the demo's scene
// (SetClothShape_Square, main.cpp:31-106 / 31-103), run the init sequence (:561-583 / :531-550: BVH topology, aEdge,
// matrix pattern) through the library's own set-up entry points (clothgpu_bvh_build_topdown,
// clothgpu_build_adjacency), step StepTime() (:213-263 / :205-237) N times with the state resident on the GPU, print
// the lines the reference prints to stdout per step (solve_internal_sparse.h:321, self_collision_cloth.cpp:386,415,
// 448-450, main.cpp:246) from the step statistics, and dump / restore the state (aXYZ, aUVW) and the vertex normals
// (MakeNormal, main.cpp:183-211) for a renderer.  The reference's "energy : W" line is not reproduced: its bending
// part is accumulated into an uninitialised variable there (solve_internal_sparse.h:65).
//
//   clothsim self_contact [--nx 4] [--nz 50] [--h 0.1] [--collider plane|sphere] ...
//   clothsim internal     [--ndiv 25] [--size 1.0] [--collider none|plane|sphere] [--solver ilu|cg] ...
//   common: --steps N  --restore FILE  --dump FILE  --dump-every K  --normals FILE  --quiet  --device D
//           --bvh-jitter (build the BVH topology on the seeded-jitter copy of the rest shape)
// state file: 8 bytes "CLOTHSIM", int32 version 1, int32 np, int32 step, np*3 doubles aXYZ, np*3 doubles aUVW.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/clothgpu.h"

namespace {

struct Scene {
  int np = 0;
  std::vector<double> xyz0;
  std::vector<int> bc, tri, quad;
  double total_area = 0;
};

// (ndiv_a+1) x (ndiv_b+1) vertices, vertex id = ia*(ndiv_b+1)+ib, two triangles per cell, one hinge per interior edge;
// plane: 0 = (x, z) with y = 0 [self_contact_sparse], 1 = (x, y) with z = 0 and the row ia = 0 pinned [internal_cloth]
Scene make_grid(int ndiv_a, int ndiv_b, double ha, double hb, int plane) {
  Scene s;
  const int nb1 = ndiv_b + 1;
  s.np = (ndiv_a + 1) * nb1;
  s.xyz0.assign((size_t)s.np * 3, 0.0);
  s.bc.assign(s.np, 0);
  for (int ia = 0; ia <= ndiv_a; ia++)
    for (int ib = 0; ib < nb1; ib++) {
      double* p = &s.xyz0[(size_t)(ia * nb1 + ib) * 3];
      p[0] = ia * ha;
      p[plane == 0 ? 2 : 1] = ib * hb;
      if (plane == 1 && ia == 0) s.bc[ia * nb1 + ib] = 1;
    }
  auto id = [nb1](int ia, int ib) { return ia * nb1 + ib; };
  for (int ia = 0; ia < ndiv_a; ia++)
    for (int ib = 0; ib < ndiv_b; ib++) {
      const int t[6] = {id(ia, ib), id(ia + 1, ib), id(ia, ib + 1), id(ia + 1, ib + 1), id(ia, ib + 1), id(ia + 1, ib)};
      s.tri.insert(s.tri.end(), t, t + 6);
    }
  for (int ia = 0; ia < ndiv_a; ia++)
    for (int ib = 0; ib < ndiv_b; ib++) {
      const int q[4] = {id(ia, ib), id(ia + 1, ib + 1), id(ia + 1, ib), id(ia, ib + 1)};
      s.quad.insert(s.quad.end(), q, q + 4);
    }
  for (int ia = 0; ia < ndiv_a; ia++)
    for (int ib = 0; ib + 1 < ndiv_b; ib++) {
      const int q[4] = {id(ia + 1, ib), id(ia, ib + 2), id(ia + 1, ib + 1), id(ia, ib + 1)};
      s.quad.insert(s.quad.end(), q, q + 4);
    }
  for (int ia = 0; ia + 1 < ndiv_a; ia++)
    for (int ib = 0; ib < ndiv_b; ib++) {
      const int q[4] = {id(ia, ib + 1), id(ia + 2, ib), id(ia + 1, ib), id(ia + 1, ib + 1)};
      s.quad.insert(s.quad.end(), q, q + 4);
    }
  s.total_area = (ndiv_a * ha) * (ndiv_b * hb);
  return s;
}

// in-plane jitter of the rest shape (xorshift64, amplitude 0.2 h): only feeds the BVH topology builder
std::vector<double> jittered(const Scene& s, double h) {
  std::vector<double> x = s.xyz0;
  bool planar[3] = {false, false, false};
  for (int a = 0; a < 3; a++)
    for (int i = 1; i < s.np; i++)
      if (x[(size_t)i * 3 + a] != x[a]) { planar[a] = true; break; }
  uint64_t st = 88172645463325252ULL;
  for (int i = 0; i < s.np; i++)
    for (int a = 0; a < 3; a++) {
      if (!planar[a]) continue;
      st ^= st << 13; st ^= st >> 7; st ^= st << 17;
      x[(size_t)i * 3 + a] += 0.2 * h * ((double)(st >> 11) / 9007199254740992.0 * 2.0 - 1.0);
    }
  return x;
}

int fail(const char* what) {
  fprintf(stderr, "clothsim: %s: %s\n", what, clothgpu_last_error());
  return 1;
}

bool write_state(const std::string& path, int np, int step, const std::vector<double>& x, const std::vector<double>& v) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  const int32_t hdr[3] = {1, np, step};
  fwrite("CLOTHSIM", 1, 8, f);
  fwrite(hdr, sizeof(int32_t), 3, f);
  fwrite(x.data(), sizeof(double), x.size(), f);
  fwrite(v.data(), sizeof(double), v.size(), f);
  return fclose(f) == 0;
}
bool read_state(const std::string& path, int np, int* step, std::vector<double>& x, std::vector<double>& v) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  char magic[8];
  int32_t hdr[3];
  bool ok = fread(magic, 1, 8, f) == 8 && !memcmp(magic, "CLOTHSIM", 8) && fread(hdr, sizeof(int32_t), 3, f) == 3 &&
            hdr[0] == 1 && hdr[1] == np;
  x.resize((size_t)np * 3);
  v.resize((size_t)np * 3);
  ok = ok && fread(x.data(), sizeof(double), x.size(), f) == x.size() && fread(v.data(), sizeof(double), v.size(), f) == v.size();
  fclose(f);
  if (ok) *step = hdr[2];
  return ok;
}

void usage() {
  fprintf(stderr,
          "usage: clothsim self_contact|internal [--nx N --nz N --h H | --ndiv N --size S] [--collider NAME]\n"
          "       [--solver ilu|cg] [--steps N] [--restore FILE] [--dump FILE] [--dump-every K] [--normals FILE]\n"
          "       [--bvh-jitter] [--device D] [--quiet]\n");
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) { usage(); return 2; }
  const std::string demo = argv[1];
  const bool self_contact = demo == "self_contact";
  if (!self_contact && demo != "internal") { usage(); return 2; }
  int nx = 4, nz = 50, ndiv = 25, steps = 10, dump_every = 0, device = -1;
  double h = 0.1, size = 1.0;
  std::string collider = self_contact ? "plane" : "none", solver = "ilu", restore, dump, normals;
  bool quiet = false, jitter = false;
  for (int i = 2; i < argc; i++) {
    const std::string a = argv[i];
    auto next = [&]() -> const char* { if (i + 1 >= argc) { usage(); exit(2); } return argv[++i]; };
    if (a == "--nx") nx = atoi(next());
    else if (a == "--nz") nz = atoi(next());
    else if (a == "--h") h = atof(next());
    else if (a == "--ndiv") ndiv = atoi(next());
    else if (a == "--size") size = atof(next());
    else if (a == "--collider") collider = next();
    else if (a == "--solver") solver = next();
    else if (a == "--steps") steps = atoi(next());
    else if (a == "--restore") restore = next();
    else if (a == "--dump") dump = next();
    else if (a == "--dump-every") dump_every = atoi(next());
    else if (a == "--normals") normals = next();
    else if (a == "--device") device = atoi(next());
    else if (a == "--quiet") quiet = true;
    else if (a == "--bvh-jitter") jitter = true;
    else { usage(); return 2; }
  }
  if (nx < 1 || nz < 1 || ndiv < 1 || steps < 0) { usage(); return 2; }

  // ---- scene and constants of the demo (self_contact_sparse/main.cpp:138-156, internal_cloth_sparse/main.cpp:137-154)
  clothgpu_params p;
  memset(&p, 0, sizeof(p));
  Scene sc;
  double hh;
  if (self_contact) {
    sc = make_grid(nx, nz, h, h, 0);
    hh = h;
    p.dt = 0.01; p.lambda = 0.0; p.myu = 30.0; p.stiff_bend = 1.0e-2;
    p.gravity[0] = 0; p.gravity[1] = 0.05; p.gravity[2] = -10;
    p.mass_point = sc.total_area * 0.04 / (double)sc.np;
    p.stiff_contact = 300; p.contact_clearance = 0.01;
    if (collider == "plane") { p.collider_kind = CLOTHGPU_COLLIDER_PLANE_Z; p.collider_par[0] = -0.5; }
    else if (collider == "sphere") {  // the container sphere of main.cpp:117-133
      p.collider_kind = CLOTHGPU_COLLIDER_SPHERE_INSIDE;
      p.collider_par[0] = 0.5; p.collider_par[1] = 0.0; p.collider_par[2] = 2.0; p.collider_par[3] = 3.0;
    } else { usage(); return 2; }
  } else {
    hh = size / ndiv;
    sc = make_grid(ndiv, ndiv, hh, hh, 1);
    p.dt = 0.02; p.lambda = 1.0; p.myu = 4.0; p.stiff_bend = 1.0e-3;
    p.gravity[2] = -10;
    p.mass_point = size * size * 1.0 / (double)sc.np;
    p.stiff_contact = 1.0e3; p.contact_clearance = 0.02;
    if (collider == "none") p.collider_kind = CLOTHGPU_COLLIDER_NONE;
    else if (collider == "plane") { p.collider_kind = CLOTHGPU_COLLIDER_PLANE_Z; p.collider_par[0] = -0.5; }
    else if (collider == "sphere") {
      p.collider_kind = CLOTHGPU_COLLIDER_SPHERE;
      p.collider_par[0] = 0.1; p.collider_par[1] = 0.5; p.collider_par[2] = -0.8; p.collider_par[3] = 0.3;
    } else { usage(); return 2; }
  }
  const int np = sc.np, ntri = (int)sc.tri.size() / 3, nquad = (int)sc.quad.size() / 4;

  // ---- init sequence: pattern (hinge quads), aEdge (triangles), BVH topology -- the library's set-up entry points
  if (device >= 0) { /* clothgpu_create selects it */ }
  std::vector<int> rowptr(np + 1), colidx, eidx(np + 1), earr, nodes;
  int n = clothgpu_build_adjacency(nquad, 4, sc.quad.data(), np, 0, rowptr.data(), nullptr, 0);
  if (n < 0) return fail("clothgpu_build_adjacency");
  colidx.resize(n > 0 ? n : 1);
  if (clothgpu_build_adjacency(nquad, 4, sc.quad.data(), np, 0, rowptr.data(), colidx.data(), (int)colidx.size()) < 0) return fail("clothgpu_build_adjacency");
  int nnode = 0;
  if (self_contact) {
    n = clothgpu_build_adjacency(ntri, 3, sc.tri.data(), np, 0, eidx.data(), nullptr, 0);
    if (n < 0) return fail("clothgpu_build_adjacency");
    earr.resize(n > 0 ? n : 1);
    if (clothgpu_build_adjacency(ntri, 3, sc.tri.data(), np, 0, eidx.data(), earr.data(), (int)earr.size()) < 0) return fail("clothgpu_build_adjacency");
    nodes.resize((size_t)(2 * ntri - 1) * 3);
    const std::vector<double> xb = jitter ? jittered(sc, hh) : sc.xyz0;
    int nfb = 0;
    nnode = clothgpu_bvh_build_topdown(ntri, sc.tri.data(), np, xb.data(), nodes.data(), 2 * ntri - 1, &nfb);
    if (nnode < 0) return fail("clothgpu_bvh_build_topdown");
    if (nfb && !quiet) printf("(BVH: %d splits where the reference's builder has no result; try --bvh-jitter)\n", nfb);
  }
  clothgpu_ctx* ctx = nullptr;
  if (clothgpu_create(&ctx, device, 1, np, ntri, nquad, sc.tri.data(), sc.quad.data(), sc.bc.data(), sc.xyz0.data(),
                      rowptr.data(), colidx.data(), nnode, nnode ? nodes.data() : nullptr, nnode ? eidx.data() : nullptr,
                      nnode ? earr.data() : nullptr) != CLOTHGPU_OK)
    return fail("clothgpu_create");

  std::vector<double> x((size_t)np * 3), v((size_t)np * 3, 0.0);
  int step0 = 0;
  if (!restore.empty()) {
    if (!read_state(restore, np, &step0, x, v)) { fprintf(stderr, "clothsim: cannot restore %s (wrong mesh size?)\n", restore.c_str()); return 1; }
    if (clothgpu_upload_state(ctx, 0, x.data(), v.data()) != CLOTHGPU_OK) return fail("clothgpu_upload_state");
  }
  if (!quiet)
    printf("clothsim %s: np=%d ntri=%d nquad=%d ncrs=%d nbvh=%d collider=%s solver=%s start step=%d\n", demo.c_str(), np,
           ntri, nquad, rowptr[np], nnode, collider.c_str(), solver.c_str(), step0);

  std::vector<int> stats(CLOTHGPU_NSTATS, 0);
  for (int s = 1; s <= steps; s++) {
    if (self_contact) {
      if (clothgpu_step_self_contact(ctx, &p, 1, stats.data()) != CLOTHGPU_OK) return fail("clothgpu_step_self_contact");
      if (!quiet) {  // the reference's stdout of one StepTime()
        double cr = 0;
        clothgpu_last_solver(ctx, &cr, nullptr);
        printf("  conv_ratio:%g  iteration:%d\n", cr, stats[0]);
        if (p.contact_clearance * 0.5 > 0) printf("  Proximity      Contact Elem Size: %d\n", stats[1]);
        for (int k = 0; k < stats[2]; k++) printf("  CCD iter: %d    Contact Elem Size: %d\n", k, stats[8 + k]);
        for (int k = 0; k < stats[3]; k++)
          printf("  RIZ iter: %d    Contact Elem Size: %d   NNode In RIZ: %d\n", k, stats[16 + k], stats[128 + k]);
        if (stats[6]) printf("Resolved All Collisions : \n");
        if (stats[4]) printf("update middle velocity\n");
      }
    } else {
      double cr = 0;
      int it = 0;
      const int rc = solver == "cg" ? clothgpu_step_internal_cg(ctx, &p, &cr, &it) : clothgpu_step_internal_ilu(ctx, &p, &cr, &it);
      if (rc != CLOTHGPU_OK) return fail("clothgpu_step_internal");
      if (!quiet) printf("  conv_ratio:%g  iteration:%d\n", cr, it);
    }
    if (dump_every > 0 && !dump.empty() && s % dump_every == 0 && s != steps) {
      if (clothgpu_download_state(ctx, 0, x.data(), v.data()) != CLOTHGPU_OK) return fail("clothgpu_download_state");
      char name[1024];
      snprintf(name, sizeof(name), "%s.%06d", dump.c_str(), step0 + s);
      if (!write_state(name, np, step0 + s, x, v)) { fprintf(stderr, "clothsim: cannot write %s\n", name); return 1; }
    }
  }
  if (clothgpu_download_state(ctx, 0, x.data(), v.data()) != CLOTHGPU_OK) return fail("clothgpu_download_state");
  double chk = 0;
  for (size_t i = 0; i < x.size(); i++) chk += x[i] * (1.0 + (double)(i % 7));
  printf("steps=%d checksum=%.17g\n", step0 + steps, chk);
  if (!dump.empty() && !write_state(dump, np, step0 + steps, x, v)) { fprintf(stderr, "clothsim: cannot write %s\n", dump.c_str()); return 1; }
  if (!normals.empty()) {
    std::vector<double> nrm((size_t)np * 3);
    if (clothgpu_make_normal(ctx, CLOTHGPU_BUF_XYZ, 0, nrm.data()) != CLOTHGPU_OK) return fail("clothgpu_make_normal");
    FILE* f = fopen(normals.c_str(), "wb");
    if (!f) { fprintf(stderr, "clothsim: cannot write %s\n", normals.c_str()); return 1; }
    fwrite(nrm.data(), sizeof(double), nrm.size(), f);
    fclose(f);
  }
  clothgpu_destroy(ctx);
  return 0;
}
