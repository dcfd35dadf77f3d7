/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may
 * load this.  The product (envpool_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED for classic_control + toy_text -- checked bit-for-bit against
 * the reference itself compiled here (oracle/_ref, see oracle/Makefile) and against
 * the committed fixtures in tests/golden/ generated from it.
 * HalfCheetah (mjc_oracle.c): PARITY UNPINNED -- MuJoCo 3.6.0 is absent (see that
 * file's header).
 */
#ifndef EP_ORACLE_H_
#define EP_ORACLE_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum epo_kind {
  EPO_CARTPOLE = 0,
  EPO_PENDULUM = 1,
  EPO_ACROBOT = 2,
  EPO_MOUNTAIN_CAR = 3,
  EPO_MOUNTAIN_CAR_CONTINUOUS = 4,
  EPO_FROZEN_LAKE = 5,
  EPO_CATCH = 6,
  EPO_TAXI = 7,
  EPO_NCHAIN = 8,
  EPO_CLIFF_WALKING = 9,
  EPO_BLACKJACK = 10,
  EPO_HALF_CHEETAH = 11,
  EPO_NUM_KINDS = 12
};

typedef struct epo_pool epo_pool;

/* iopt: FrozenLake size / Pendulum version / CliffWalking is_slippery /
 * Blackjack (natural | sab<<1); -1 = reference default.
 * env_seed: optional per-env seeds (envpool/core/env.h:101-111), else seed+env_id.
 * max_episode_steps <= 0 selects INT_MAX (envpool/core/env_spec.h:31). */
epo_pool* epo_create(int kind, int num_envs, int seed, const int* env_seed,
                     int max_episode_steps, int iopt);
void epo_destroy(epo_pool* p);
/* Forced reset of the listed envs (AsyncEnvPool::Reset, async_envpool.h:224-237);
 * env_ids NULL = all envs in order.  Output row i <-> env_ids[i]. */
void epo_reset(epo_pool* p, const int32_t* env_ids, int n);
/* One Send+Recv in sync mode (async_envpool.h:59-82,118-131,169-181). */
void epo_step(epo_pool* p, const void* action, const int32_t* env_ids, int n);
int epo_num_keys(const epo_pool* p);
const char* epo_key_name(const epo_pool* p, int k);
/* element size in bytes and per-row element count of state key k */
int epo_key_elem_size(const epo_pool* p, int k);
int epo_key_row_elems(const epo_pool* p, int k);
/* data of state key k for the rows of the last reset/step call */
const void* epo_key_data(const epo_pool* p, int k);
int epo_action_elem_size(const epo_pool* p);
int epo_action_row_elems(const epo_pool* p);
/* teacher forcing (tests): read / overwrite one env's continuous state s[0..4], its
 * done flag and its step counters (current_step_ == elapsed_step_ for every env that has
 * one) without touching its RNG */
void epo_get_state(const epo_pool* p, int eid, double* s5, int* done, int* cur);
void epo_set_state(epo_pool* p, int eid, const double* s5, int done, int cur);
/* HalfCheetah teacher forcing: s27 = qpos[9] qvel[9] qacc_warmstart[9] */
void epo_mjc_set(epo_pool* p, int eid, const double* s27, int done, int cur);
void epo_mjc_get(const epo_pool* p, int eid, double* s27);
/* raw engine draw from env `eid`'s std::mt19937 (for RNG known-answer tests) */
uint32_t epo_debug_draw(epo_pool* p, int eid);
void epo_debug_set_rng(epo_pool* p, int eid, const uint32_t* mt624, int idx);
int epo_debug_uniform_int(epo_pool* p, int eid, int a, int b);
double epo_debug_uniform_real(epo_pool* p, int eid, double a, double b);
double epo_debug_normal(epo_pool* p, int eid, double mean, double stddev);

#ifdef __cplusplus
}
#endif
#endif /* EP_ORACLE_H_ */
