/* oracle/cloth_oracle.h -- flat C interface of the CPU restatement (TEST INFRASTRUCTURE ONLY).
 *
 * Every orc_* function has a ref_* twin with the same signature in oracle/ref_shim.cpp (the unmodified reference
 * behind a C wrapper); oracle/oracle_py.py binds both through one code path.  Array conventions:
 *   xyz, uvw, xyz0 : np x 3 doubles          tri : ntri x 3 ints        quad : nquad x 4 ints
 *   rowptr[np+1], colidx[ncrs]               : block-CRS pattern, columns ascending per row, no diagonal
 *   valDia : np x 9, valCrs : ncrs x 9       : row-major 3x3 blocks
 *   nodes  : nnode x {parent, child0, child1}; a leaf has child1 == -1 and child0 == triangle index
 *   boxes  : nnode x {xmin,xmax,ymin,ymax,zmin,zmax}
 *   elems  : n x {ino0,ino1,ino2,ino3,is_fv} in the reference's std::set order
 */
#ifndef CLOTH_ORACLE_H
#define CLOTH_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_NSTATS = 128 };

int orc_edge_of_elem(const int* elem, int nelem, int nnoel, int nnode, int do_sort, int* index, int* array, int cap);
void orc_ilu_diaind(int np, const int* rowptr, const int* colidx, int* diaind);

void orc_cst(const double* C, const double* c, double lambda, double myu, double* W, double* dW, double* ddW);
void orc_bend(const double* C, const double* c, double stiff, double* W, double* dW, double* ddW);
void orc_contact(const double* c, double stiff, double clearance, int col_kind, const double* col_par, double* W,
                 double* dW, double* ddW);

void orc_assemble(int np, int ntri, int nquad, const int* tri, const int* quad, const int* bc, const double* xyz0,
                  const double* xyz, const double* uvw, const int* rowptr, const int* colidx, double dt, double lambda,
                  double myu, double stiff_bend, const double* gravity, double mass_point, double stiff_contact,
                  double clearance, int col_kind, const double* col_par, double* valDia, double* valCrs, double* vec_b);
void orc_matvec(int np, const int* rowptr, const int* colidx, const double* valDia, const double* valCrs, double alpha,
                const double* x, double beta, double* y);
double orc_dot(int n, const double* a, const double* b);
void orc_ilu_factor(int np, const int* rowptr, const int* colidx, double* valDia, double* valCrs);
void orc_ilu_solve(int np, const int* rowptr, const int* colidx, const double* iluDia, const double* iluCrs, double* vec);
void orc_pcg(int np, const int* rowptr, const int* colidx, const double* valDia, const double* valCrs,
             const double* iluDia, const double* iluCrs, double* r, double* x, double* conv_ratio, int* iteration);
void orc_cg(int np, const int* rowptr, const int* colidx, const double* valDia, const double* valCrs, double* r,
            double* x, double* conv_ratio, int* iteration);
void orc_make_normal(int np, int ntri, const int* tri, const double* xyz, double* normal);
void orc_update_intermediate_velocity(int np, int ntri, int nquad, const int* tri, const int* quad, const int* bc,
                                      const double* xyz0, const double* xyz, double* uvw, double dt, double lambda,
                                      double myu, double stiff_bend, const double* gravity, double mass_point);

void orc_refit_prx(int nnode, const int* nodes, int ntri, const int* tri, int np, const double* xyz, double delta,
                   double* boxes);
void orc_refit_ccd(int nnode, const int* nodes, int ntri, const int* tri, int np, const double* xyz, const double* uvw,
                   double dt, double* boxes);
int orc_leaf_pairs(int nnode, const int* nodes, const double* boxes, int* pairs, int cap);
int orc_contacts_prx(int nnode, const int* nodes, const double* boxes, int ntri, const int* tri, int np,
                     const double* xyz, double delta, int* elems, int cap);
int orc_contacts_ccd(int nnode, const int* nodes, const double* boxes, int ntri, const int* tri, int np,
                     const double* xyz, const double* uvw, double dt, double delta, int* elems, int cap);

double orc_dist_fv(const double* p, double* w0, double* w1);
double orc_dist_ee(const double* p, double* rp, double* rq);
double orc_coplanar(const double* s, const double* e);
int orc_is_fv_prx(const int* id, const double* p, const double* box, double delta);
int orc_is_ee_prx(const int* id, const double* p, double delta);
int orc_is_fv_ccd(const int* id, const double* p, const double* q, const double* box);
int orc_is_ee_ccd(const int* id, const double* p, const double* q);

void orc_impulse_prx(int np, double* uvw, double delta, double stiffness, double dt, double mass, const double* xyz,
                     int nelem, const int* elems);
void orc_impulse_ccd(int np, double* uvw, double delta, double stiffness, double dt, double mass, const double* xyz,
                     int nelem, const int* elems);

void* orc_riz_new(void);
void orc_riz_free(void* h);
void orc_riz_make(void* h, int nelem, const int* elems, int np, const int* edge_index, const int* edge_array);
int orc_riz_get(void* h, int np, int* zone_of_node);
void orc_riz_apply(void* h, int np, double* uvw, const double* xyz, const double* uvw0);

void* orc_sim_new(int np, int ntri, int nquad, const int* tri, const int* quad, const int* bc, const double* xyz0,
                  int nnode, const int* nodes);
void orc_sim_free(void* h);
void orc_sim_set_state(void* h, const double* xyz, const double* uvw);
void orc_sim_get_state(void* h, double* xyz, double* uvw);
void orc_sim_step_internal(void* h, double dt, double lambda, double myu, double stiff_bend, const double* gravity,
                           double mass_point, double stiff_contact, double clearance, int col_kind,
                           const double* col_par, double* conv_ratio, int* iteration);
void orc_sim_step_internal_cg(void* h, double dt, double lambda, double myu, double stiff_bend, const double* gravity,
                           double mass_point, double stiff_contact, double clearance, int col_kind,
                           const double* col_par, double* conv_ratio, int* iteration);
int orc_sim_resolve_contacts(void* h, double* uvw, const double* xyz_prev, double dt, double clearance,
                             double mass_point, double stiffness, int* stats);
void orc_sim_step_self_contact(void* h, double dt, double lambda, double myu, double stiff_bend, const double* gravity,
                               double mass_point, double stiff_contact, double clearance, int col_kind,
                               const double* col_par, int* stats);

#ifdef __cplusplus
}
#endif
#endif
