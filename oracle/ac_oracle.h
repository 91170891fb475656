/*
 * ac_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's AC-move transition function and of
 * the two searches built on it.  It exists only so that tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() can check the
 * CUDA path bit for bit.  Nothing under ac_solver_b200/ may include, link or
 * call it.
 *
 * Parity status: PINNED.  oracle/gen_golden.py imports the real reference from
 * /root/reference in the build container and writes tests/golden/ ; the tests in
 * tests/test_oracle_golden.py replay every fixture through this library
 * (unit vectors of the reference's tests/test_ac_env.py, the known-answer paths
 * of tests/search/, the shipped Miller-Schupp greedy paths, random differential
 * triples including the AssertionError / IndexError cases).
 *
 * Reference citations are relative to /root/reference.
 */
#ifndef AC_ORACLE_H
#define AC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-call / per-row status: what the reference would have done. */
#define ACO_OK 0     /* returned normally                                              */
#define ACO_ASSERT 1 /* AssertionError: is_array_valid_presentation failed, utils.py:261 */
#define ACO_INDEX 2  /* IndexError: conjugate() indexed an empty relator, ac_moves.py:119 */

/* utils.py:175-240  simplify_relator on the first n (non-zero) letters, in place.
 * Works for any int8 alphabet.  Returns the reduced length (letters beyond it are
 * left untouched; callers pad). */
int aco_simplify_relator(int8_t *rel, int n, int cyclical);

/* utils.py:13-54  is_array_valid_presentation for an array of 2*mrl letters. */
int aco_is_valid_presentation(const int8_t *p, int mrl);

/* ac_moves.py:4-76 / :79-156  the raw moves, WITHOUT simplify_presentation, in place.
 * Return the new length of r_i, -1 if rejected (array unchanged), -2 if the reference
 * raises IndexError, -3 if it raises AssertionError on the arguments. */
int aco_concatenate_relators(int8_t *p, int mrl, int i, int j, int sign);
int aco_conjugate(int8_t *p, int mrl, int i, int j, int sign);

/* ac_moves.py:159-231  ACMove(move_id, presentation, mrl, lengths(ignored), cyclical).
 * in/out are 2*mrl letters (may alias).  lens_out[2] receives the recomputed
 * lengths.  Returns ACO_*; on a non-OK status out/lens_out are unspecified. */
int aco_acmove(int move_id, const int8_t *in, int mrl, int cyclical, int8_t *out,
               int *lens_out);

/* Batched ACMove over N rows; row stride 2*mrl.  nthreads <= 0: all cores. */
void aco_moves_batch(const int8_t *in, const uint8_t *action, int8_t *out,
                     uint8_t *lens_out /*[N,2]*/, uint8_t *status /*[N]*/, int64_t n,
                     int mrl, int cyclical, int nthreads);

/* ac_env.py:95-113  ACEnv.step over N independent envs, state updated in place.
 * reward = horizon*mrl*2 if done else -(len0+len1); step_count += 1;
 * truncated = step_count >= horizon.  Rows with a non-OK status keep their state. */
void aco_env_step_batch(int8_t *state, const uint8_t *action, int32_t *reward,
                        uint8_t *done, uint8_t *truncated, int32_t *step_count,
                        uint8_t *lens /*[N,2]*/, uint8_t *status, int64_t n, int mrl,
                        int horizon, int nthreads);

/* Search result block shared by bfs and greedy. */
typedef struct {
    int32_t solved;          /* 1 = a child of total length 2 was generated             */
    int32_t status;          /* ACO_* of the move that raised, or ACO_OK                 */
    int32_t budget_hit;      /* 1 = the "Exiting search ..." line would be printed       */
    int32_t path_len;        /* entries written to path (action,length pairs)           */
    int64_t n_visited;       /* len(tree_nodes) at return                               */
    int64_t n_expanded;      /* nodes popped and expanded (12 children generated)       */
    int64_t n_moves;         /* ACMove calls                                            */
    int64_t frontier_left;   /* len(to_explore) at return                               */
    int32_t n_minlen;        /* entries in minlen_log                                   */
    int32_t minlen_log[128]; /* successive "New minimal length found" values            */
} aco_search_result;

/* breadth_first.py:15-97  bfs().  path: int32 pairs (action, total_length), capacity
 * path_cap pairs.  visited_out (optional): receives the visited states in insertion
 * order, up to visited_cap rows of 2*mrl letters. */
int aco_bfs(const int8_t *presentation, int mrl, int64_t max_nodes, int cyclical,
            int32_t *path, int path_cap, int8_t *visited_out, int64_t visited_cap,
            aco_search_result *res);

/* greedy.py:15-121  greedy_search().  Same conventions; on failure path holds the
 * reference's (False, path + [(11, len)]) value. */
int aco_greedy(const int8_t *presentation, int mrl, int64_t max_nodes, int cyclical,
               int32_t *path, int path_cap, int8_t *visited_out, int64_t visited_cap,
               aco_search_result *res);

int aco_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
