/* placeholder -- replaced by the full restatement */
#include "mjc_oracle.h"
#include <stdlib.h>
struct mjc_model { int dummy; };
struct mjc_data { double qpos[9], qvel[9]; };
mjc_model* mjc_make_half_cheetah(void) { return (mjc_model*)calloc(1, sizeof(mjc_model)); }
void mjc_free_model(mjc_model* m) { free(m); }
mjc_data* mjc_make_data(const mjc_model* m) { (void)m; return (mjc_data*)calloc(1, sizeof(mjc_data)); }
void mjc_free_data(mjc_data* d) { free(d); }
void mjc_reset(const mjc_model* m, mjc_data* d, double s, mjc_uniform_fn u, mjc_normal_fn n, void* c) { (void)m;(void)d;(void)s;(void)u;(void)n;(void)c; abort(); }
void mjc_step(const mjc_model* m, mjc_data* d, const double* a, int f) { (void)m;(void)d;(void)a;(void)f; abort(); }
const double* mjc_qpos(const mjc_data* d) { return d->qpos; }
const double* mjc_qvel(const mjc_data* d) { return d->qvel; }
double* mjc_qpos_mut(mjc_data* d) { return d->qpos; }
double* mjc_qvel_mut(mjc_data* d) { return d->qvel; }
void mjc_forward(const mjc_model* m, mjc_data* d) { (void)m;(void)d; abort(); }
int mjc_nefc(const mjc_data* d) { (void)d; return 0; }
int mjc_model_constants(const mjc_model* m, double* out, int cap) { (void)m;(void)out;(void)cap; return 0; }
