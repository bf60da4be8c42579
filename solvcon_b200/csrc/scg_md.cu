// scg_md.cu -- multidim Euler CESE on the device (placeholder while the kernels land).
#include "scg_common.cuh"
using namespace scg;
#define NOTYET(name) return fail(name ": not implemented yet")
extern "C" {
int scg_mesh_create(scg_mesh **, int, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, const int32_t *, const int32_t *, const int32_t *, const int32_t *, const double *, const double *, const double *, const double *, int) { NOTYET("scg_mesh_create"); }
int scg_mesh_destroy(scg_mesh *) { return 0; }
int scg_mesh_perm(scg_mesh *, int32_t *) { NOTYET("scg_mesh_perm"); }
int scg_md_create(scg_md **, scg_mesh *, double) { NOTYET("scg_md_create"); }
int scg_md_destroy(scg_md *) { return 0; }
int scg_md_set_params(scg_md *, double, double, double) { NOTYET("scg_md_set_params"); }
int scg_md_init_solution(scg_md *, double, double, const double *, double) { NOTYET("scg_md_init_solution"); }
int scg_md_add_bc(scg_md *, int, const int32_t *, int64_t, const double *, int) { NOTYET("scg_md_add_bc"); }
int scg_md_clear_bc(scg_md *) { NOTYET("scg_md_clear_bc"); }
int scg_md_upload(scg_md *, int, const double *) { NOTYET("scg_md_upload"); }
int scg_md_download(scg_md *, int, double *) { NOTYET("scg_md_download"); }
int scg_md_phase(scg_md *, int) { NOTYET("scg_md_phase"); }
int scg_md_march(scg_md *, int64_t) { NOTYET("scg_md_march"); }
int scg_md_reduce(scg_md *, const double *, double *, double *, double *) { NOTYET("scg_md_reduce"); }
int scg_md_algorithmic_bytes(scg_md *, int64_t *) { NOTYET("scg_md_algorithmic_bytes"); }
int scg_md_normal_matrix(scg_md *, int32_t, double *) { NOTYET("scg_md_normal_matrix"); }
int scg_md_halo_set(scg_md *, const int32_t *, int64_t, const int32_t *, int64_t) { NOTYET("scg_md_halo_set"); }
int scg_md_halo_pack(scg_md *, double **, int64_t *) { NOTYET("scg_md_halo_pack"); }
int scg_md_halo_unpack(scg_md *, const double *) { NOTYET("scg_md_halo_unpack"); }
}
