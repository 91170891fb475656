/*
 * ac_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See ac_oracle.h for the contract and the parity status (PINNED against the
 * reference through tests/golden/).  Citations are relative to /root/reference.
 *
 * The reference works on numpy arrays with Python containers; this file restates
 * the same array semantics with plain C arrays: "copy, filter the non-zero
 * letters, rewrite one half, validate, reduce both halves, pad".
 */
#include "ac_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ACO_MAX_MRL 127

int aco_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* Word arithmetic                                                            */
/* ------------------------------------------------------------------------- */

static int count_nonzero(const int8_t *p, int n) {
    int c = 0;
    for (int t = 0; t < n; ++t) c += (p[t] != 0);
    return c;
}

/* gather the non-zero letters of p[0..n) in order (numpy boolean-mask semantics,
 * ac_moves.py:50-51,109) */
static int gather_nonzero(const int8_t *p, int n, int8_t *dst) {
    int c = 0;
    for (int t = 0; t < n; ++t)
        if (p[t] != 0) dst[c++] = p[t];
    return c;
}

/* utils.py:207-217: scan with back-tracking deletes adjacent inverse pairs until
 * none is left.  The freely reduced form of a word is unique, so a stack gives the
 * same array.  utils.py:220-229: with cyclical, strip pos matching outer pairs. */
int aco_simplify_relator(int8_t *rel, int n, int cyclical) {
    int top = 0; /* rel[0..top) is the reduced prefix */
    for (int t = 0; t < n; ++t) {
        int8_t c = rel[t];
        if (top > 0 && rel[top - 1] == (int8_t)(-c))
            --top;
        else
            rel[top++] = c;
    }
    n = top;
    if (cyclical && n > 0) {
        int pos = 0;
        /* terminates before the middle for a reduced word of non-zero letters */
        while (pos < n && rel[pos] == (int8_t)(-rel[n - pos - 1])) ++pos;
        if (pos) {
            memmove(rel, rel + pos, (size_t)(n - 2 * pos));
            n -= 2 * pos;
        }
    }
    return n;
}

/* utils.py:13-54 */
int aco_is_valid_presentation(const int8_t *p, int mrl) {
    int l0 = count_nonzero(p, mrl);
    int l1 = count_nonzero(p + mrl, mrl);
    if (l0 == 0 || l1 == 0) return 0;
    for (int t = l0; t < mrl; ++t)
        if (p[t] != 0) return 0;
    for (int t = l1; t < mrl; ++t)
        if (p[mrl + t] != 0) return 0;
    return 1;
}

/* ac_moves.py:4-76: r_i <- r_i r_j^{sign}, cancelling at the junction only; the
 * rewrite happens iff the candidate has at most mrl letters. */
static int concat_move(int8_t *P, int mrl, int i, int sign) {
    int j = 1 - i;
    int8_t u[ACO_MAX_MRL], v[ACO_MAX_MRL];
    int lu = gather_nonzero(P + i * mrl, mrl, u);
    int lv;
    if (sign == 1) {
        lv = gather_nonzero(P + j * mrl, mrl, v);
    } else { /* reversed and negated, ac_moves.py:43-48 */
        lv = 0;
        for (int t = mrl - 1; t >= 0; --t)
            if (P[j * mrl + t] != 0) v[lv++] = (int8_t)(-P[j * mrl + t]);
    }
    int acc = 0;
    int lim = lu < lv ? lu : lv;
    while (acc < lim && u[lu - 1 - acc] == (int8_t)(-v[acc])) ++acc;
    int new_size = lu + lv - 2 * acc;
    if (new_size <= mrl) {
        int8_t *dst = P + i * mrl;
        memcpy(dst, u, (size_t)(lu - acc));
        memcpy(dst + lu - acc, v + acc, (size_t)(lv - acc));
        memset(dst + new_size, 0, (size_t)(mrl - new_size));
        return new_size;
    }
    return -1; /* rejected: presentation unchanged */
}

/* ac_moves.py:79-156: r_i <- g r_i g^{-1} with at most one letter cancelled at
 * each end; literal restatement of the three slice writes. */
static int conj_move(int8_t *P, int mrl, int i, int g, int *new_size_out) {
    int8_t rel[ACO_MAX_MRL];
    int size = gather_nonzero(P + i * mrl, mrl, rel);
    if (size == 0) return ACO_INDEX; /* relator_nonzero[0], ac_moves.py:119 */
    int s = rel[0] == (int8_t)(-g);
    int e = rel[size - 1] == (int8_t)g;
    int new_size = size + 2 - 2 * (s + e);
    if (new_size <= mrl) {
        int8_t *dst = P + i * mrl;
        int n = size - e - s; /* letters kept: rel[s : size-e] */
        if (n > 0) memcpy(dst + 1 - s, rel + s, (size_t)n);
        if (!s) dst[0] = (int8_t)g;
        if (!e) dst[size + 1 - 2 * s] = (int8_t)(-g);
        if (s && e) {
            for (int t = new_size; t < new_size + 2 && t < mrl; ++t) dst[t] = 0;
        }
        *new_size_out = new_size;
    } else {
        *new_size_out = -1;
    }
    return ACO_OK;
}

/* raw moves without the trailing simplify_presentation (the reference's unit tests
 * call them directly, tests/test_ac_env.py:184-477).  Return the new length of r_i,
 * -1 if the move was rejected (array unchanged), -2 for the IndexError case. */
int aco_concatenate_relators(int8_t *P, int mrl, int i, int j, int sign) {
    if (!((i == 0 || i == 1) && j == 1 - i) || !(sign == 1 || sign == -1)) return -3;
    return concat_move(P, mrl, i, sign);
}

int aco_conjugate(int8_t *P, int mrl, int i, int j, int sign) {
    if (!((i == 0 || i == 1) && (j == 1 || j == 2)) || !(sign == 1 || sign == -1)) return -3;
    int ns = -1;
    if (conj_move(P, mrl, i, sign * j, &ns) != ACO_OK) return -2;
    return ns;
}

/* ac_moves.py:159-231 */
int aco_acmove(int move_id, const int8_t *in, int mrl, int cyclical, int8_t *out,
               int *lens_out) {
    int8_t P[2 * ACO_MAX_MRL];
    if (move_id < 0 || move_id > 11 || mrl < 1 || mrl > ACO_MAX_MRL) return ACO_ASSERT;
    memcpy(P, in, (size_t)(2 * mrl));
    int m = move_id + 1;
    int i = m % 2;
    if (move_id < 4) {
        int sign = (((m - i) / 2) % 2) ? -1 : 1; /* ac_moves.py:192-198 */
        concat_move(P, mrl, i, sign);
    } else {
        int jp = ((m - i) / 2) % 2; /* ac_moves.py:199-206 */
        int sp = ((m - i - 2 * jp) / 4) % 2;
        int g = (sp ? -1 : 1) * (jp + 1);
        int ns;
        int st = conj_move(P, mrl, i, g, &ns);
        if (st != ACO_OK) return st;
    }
    /* utils.py:243-280 simplify_presentation: validate, then reduce both halves */
    if (!aco_is_valid_presentation(P, mrl)) return ACO_ASSERT;
    for (int k = 0; k < 2; ++k) {
        int8_t *r = P + k * mrl;
        int n = count_nonzero(r, mrl);
        int nn = aco_simplify_relator(r, n, cyclical);
        memset(r + nn, 0, (size_t)(mrl - nn));
        lens_out[k] = nn;
    }
    memcpy(out, P, (size_t)(2 * mrl));
    return ACO_OK;
}

void aco_moves_batch(const int8_t *in, const uint8_t *action, int8_t *out,
                     uint8_t *lens_out, uint8_t *status, int64_t n, int mrl,
                     int cyclical, int nthreads) {
    if (nthreads <= 0) nthreads = aco_num_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t r = 0; r < n; ++r) {
        int lens[2] = {0, 0};
        int st = aco_acmove(action[r], in + r * 2 * mrl, mrl, cyclical, out + r * 2 * mrl,
                            lens);
        if (st != ACO_OK && out != in)
            memcpy(out + r * 2 * mrl, in + r * 2 * mrl, (size_t)(2 * mrl));
        if (lens_out) {
            lens_out[2 * r] = (uint8_t)lens[0];
            lens_out[2 * r + 1] = (uint8_t)lens[1];
        }
        if (status) status[r] = (uint8_t)st;
    }
}

/* ac_env.py:95-113 */
void aco_env_step_batch(int8_t *state, const uint8_t *action, int32_t *reward,
                        uint8_t *done, uint8_t *truncated, int32_t *step_count,
                        uint8_t *lens, uint8_t *status, int64_t n, int mrl, int horizon,
                        int nthreads) {
    if (nthreads <= 0) nthreads = aco_num_threads();
    const int32_t max_reward = horizon * mrl * 2; /* ac_env.py:80 */
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t r = 0; r < n; ++r) {
        int l[2] = {0, 0};
        int8_t *s = state + r * 2 * mrl;
        int st = aco_acmove(action[r], s, mrl, 1, s, l);
        if (status) status[r] = (uint8_t)st;
        if (st != ACO_OK) continue;
        int tot = l[0] + l[1];
        int d = tot == 2;
        reward[r] = d ? max_reward : -tot; /* ac_env.py:101-102 */
        done[r] = (uint8_t)d;
        step_count[r] += 1;
        truncated[r] = (uint8_t)(step_count[r] >= horizon); /* ac_env.py:104-105 */
        if (lens) {
            lens[2 * r] = (uint8_t)l[0];
            lens[2 * r + 1] = (uint8_t)l[1];
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Search substrate: node store + exact visited set                          */
/* ------------------------------------------------------------------------- */

typedef struct {
    int W;          /* letters per state = 2*mrl */
    int8_t *states; /* [cap, W] insertion order */
    int64_t *parent;
    int8_t *action;
    uint8_t *total_len;
    int32_t *depth;
    int64_t n, cap;
    int64_t *table; /* open addressing, -1 empty, else node index */
    int64_t tcap;   /* power of two */
} node_store;

static uint64_t hash_state(const int8_t *s, int W) {
    uint64_t h = 1469598103934665603ull;
    for (int t = 0; t < W; ++t) {
        h ^= (uint8_t)s[t];
        h *= 1099511628211ull;
    }
    h ^= h >> 29;
    h *= 0xbf58476d1ce4e5b9ull;
    h ^= h >> 32;
    return h;
}

static int ns_init(node_store *ns, int W) {
    memset(ns, 0, sizeof(*ns));
    ns->W = W;
    ns->cap = 1024;
    ns->states = (int8_t *)malloc((size_t)ns->cap * W);
    ns->parent = (int64_t *)malloc((size_t)ns->cap * sizeof(int64_t));
    ns->action = (int8_t *)malloc((size_t)ns->cap);
    ns->total_len = (uint8_t *)malloc((size_t)ns->cap);
    ns->depth = (int32_t *)malloc((size_t)ns->cap * sizeof(int32_t));
    ns->tcap = 4096;
    ns->table = (int64_t *)malloc((size_t)ns->tcap * sizeof(int64_t));
    if (!ns->states || !ns->parent || !ns->action || !ns->total_len || !ns->depth ||
        !ns->table)
        return -1;
    for (int64_t t = 0; t < ns->tcap; ++t) ns->table[t] = -1;
    return 0;
}

static void ns_free(node_store *ns) {
    free(ns->states);
    free(ns->parent);
    free(ns->action);
    free(ns->total_len);
    free(ns->depth);
    free(ns->table);
}

static void ns_rehash(node_store *ns) {
    int64_t ncap = ns->tcap * 2;
    int64_t *nt = (int64_t *)malloc((size_t)ncap * sizeof(int64_t));
    for (int64_t t = 0; t < ncap; ++t) nt[t] = -1;
    for (int64_t k = 0; k < ns->n; ++k) {
        uint64_t h = hash_state(ns->states + k * ns->W, ns->W) & (uint64_t)(ncap - 1);
        while (nt[h] != -1) h = (h + 1) & (uint64_t)(ncap - 1);
        nt[h] = k;
    }
    free(ns->table);
    ns->table = nt;
    ns->tcap = ncap;
}

/* returns node index if present, else -1 */
static int64_t ns_find(const node_store *ns, const int8_t *s) {
    uint64_t h = hash_state(s, ns->W) & (uint64_t)(ns->tcap - 1);
    while (ns->table[h] != -1) {
        if (memcmp(ns->states + ns->table[h] * ns->W, s, (size_t)ns->W) == 0)
            return ns->table[h];
        h = (h + 1) & (uint64_t)(ns->tcap - 1);
    }
    return -1;
}

static int64_t ns_add(node_store *ns, const int8_t *s, int64_t parent, int action,
                      int total_len, int depth) {
    if (ns->n == ns->cap) {
        ns->cap *= 2;
        ns->states = (int8_t *)realloc(ns->states, (size_t)ns->cap * ns->W);
        ns->parent = (int64_t *)realloc(ns->parent, (size_t)ns->cap * sizeof(int64_t));
        ns->action = (int8_t *)realloc(ns->action, (size_t)ns->cap);
        ns->total_len = (uint8_t *)realloc(ns->total_len, (size_t)ns->cap);
        ns->depth = (int32_t *)realloc(ns->depth, (size_t)ns->cap * sizeof(int32_t));
    }
    int64_t k = ns->n++;
    memcpy(ns->states + k * ns->W, s, (size_t)ns->W);
    ns->parent[k] = parent;
    ns->action[k] = (int8_t)action;
    ns->total_len[k] = (uint8_t)total_len;
    ns->depth[k] = depth;
    if (ns->n * 2 > ns->tcap) ns_rehash(ns);
    else {
        uint64_t h = hash_state(s, ns->W) & (uint64_t)(ns->tcap - 1);
        while (ns->table[h] != -1) h = (h + 1) & (uint64_t)(ns->tcap - 1);
        ns->table[h] = k;
    }
    return k;
}

/* path of node k from the root: [(−1,L0), (a1,L1), ...] then the extra entry */
static void write_path(const node_store *ns, int64_t k, int extra_action, int extra_len,
                       int32_t *path, int path_cap, aco_search_result *res) {
    int d = 0;
    for (int64_t q = k; q >= 0; q = ns->parent[q]) ++d;
    int total = d + 1;
    res->path_len = total;
    if (!path) return;
    int pos = d - 1;
    for (int64_t q = k; q >= 0; q = ns->parent[q], --pos) {
        if (pos < path_cap) {
            path[2 * pos] = ns->action[q];
            path[2 * pos + 1] = ns->total_len[q];
        }
    }
    if (d < path_cap) {
        path[2 * d] = extra_action;
        path[2 * d + 1] = extra_len;
    }
}

static void dump_visited(const node_store *ns, int8_t *out, int64_t cap) {
    if (!out) return;
    int64_t m = ns->n < cap ? ns->n : cap;
    memcpy(out, ns->states, (size_t)m * ns->W);
}

static void note_minlen(aco_search_result *res, int *min_length, int new_length) {
    if (new_length < *min_length) { /* breadth_first.py:79-82, greedy.py:82-85 */
        *min_length = new_length;
        if (res->n_minlen < 128) res->minlen_log[res->n_minlen++] = new_length;
    }
}

/* breadth_first.py:15-97 */
int aco_bfs(const int8_t *presentation, int mrl, int64_t max_nodes, int cyclical,
            int32_t *path, int path_cap, int8_t *visited_out, int64_t visited_cap,
            aco_search_result *res) {
    memset(res, 0, sizeof(*res));
    if (!aco_is_valid_presentation(presentation, mrl)) { /* :36-38 */
        res->status = ACO_ASSERT;
        return ACO_ASSERT;
    }
    const int W = 2 * mrl;
    node_store ns;
    if (ns_init(&ns, W)) return -1;
    int L0 = count_nonzero(presentation, mrl) + count_nonzero(presentation + mrl, mrl);
    ns_add(&ns, presentation, -1, -1, L0, 0); /* :55-58 */
    int min_length = L0;
    int64_t head = 0; /* FIFO == insertion order, so the queue is an index */
    int8_t child[2 * ACO_MAX_MRL];
    int rc = ACO_OK;
    while (head < ns.n) {
        int64_t cur = head++; /* popleft, :62 */
        res->n_expanded++;
        for (int a = 0; a < 12; ++a) { /* :69 */
            int lens[2];
            int st = aco_acmove(a, ns.states + cur * W, mrl, cyclical, child, lens);
            res->n_moves++;
            if (st != ACO_OK) {
                res->status = st;
                rc = st;
                goto out;
            }
            int nl = lens[0] + lens[1];
            note_minlen(res, &min_length, nl);
            if (nl == 2) { /* :84-85, before the visited test */
                res->solved = 1;
                write_path(&ns, cur, a, nl, path, path_cap, res);
                goto out;
            }
            if (ns_find(&ns, child) < 0) /* :87-89 */
                ns_add(&ns, child, cur, a, nl, ns.depth[cur] + 1);
        }
        if (ns.n >= max_nodes) { /* :91-95, only after all 12 children */
            res->budget_hit = 1;
            break;
        }
    }
out:
    res->n_visited = ns.n;
    res->frontier_left = ns.n - head;
    dump_visited(&ns, visited_out, visited_cap);
    ns_free(&ns);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* greedy.py:15-121: min-heap on (total_len, depth, state tuple)              */
/* ------------------------------------------------------------------------- */

static int node_less(const node_store *ns, int64_t a, int64_t b) {
    if (ns->total_len[a] != ns->total_len[b]) return ns->total_len[a] < ns->total_len[b];
    if (ns->depth[a] != ns->depth[b]) return ns->depth[a] < ns->depth[b];
    /* tuple comparison of signed letters, padding zeros included (greedy.py:104-113) */
    const int8_t *x = ns->states + a * ns->W, *y = ns->states + b * ns->W;
    for (int t = 0; t < ns->W; ++t)
        if (x[t] != y[t]) return x[t] < y[t];
    return 0;
}

typedef struct {
    int64_t *h;
    int64_t n, cap;
} heap64;

static void heap_push(heap64 *hp, const node_store *ns, int64_t k) {
    if (hp->n == hp->cap) {
        hp->cap = hp->cap ? hp->cap * 2 : 1024;
        hp->h = (int64_t *)realloc(hp->h, (size_t)hp->cap * sizeof(int64_t));
    }
    int64_t i = hp->n++;
    while (i > 0) {
        int64_t p = (i - 1) / 2;
        if (!node_less(ns, k, hp->h[p])) break;
        hp->h[i] = hp->h[p];
        i = p;
    }
    hp->h[i] = k;
}

static int64_t heap_pop(heap64 *hp, const node_store *ns) {
    int64_t top = hp->h[0];
    int64_t last = hp->h[--hp->n];
    int64_t i = 0;
    for (;;) {
        int64_t c = 2 * i + 1;
        if (c >= hp->n) break;
        if (c + 1 < hp->n && node_less(ns, hp->h[c + 1], hp->h[c])) ++c;
        if (!node_less(ns, hp->h[c], last)) break;
        hp->h[i] = hp->h[c];
        i = c;
    }
    if (hp->n > 0) hp->h[i] = last;
    return top;
}

int aco_greedy(const int8_t *presentation, int mrl, int64_t max_nodes, int cyclical,
               int32_t *path, int path_cap, int8_t *visited_out, int64_t visited_cap,
               aco_search_result *res) {
    memset(res, 0, sizeof(*res));
    if (mrl < 1 || mrl > ACO_MAX_MRL) return -1;
    const int W = 2 * mrl;
    node_store ns;
    if (ns_init(&ns, W)) return -1;
    heap64 hp = {0, 0, 0};
    int L0 = count_nonzero(presentation, mrl) + count_nonzero(presentation + mrl, mrl);
    ns_add(&ns, presentation, -1, -1, L0, 0);
    heap_push(&hp, &ns, 0);
    int min_length = L0;
    int8_t child[2 * ACO_MAX_MRL];
    int rc = ACO_OK;
    int64_t cur = 0;
    int last_len = L0;
    while (hp.n > 0) {
        cur = heap_pop(&hp, &ns); /* greedy.py:71 */
        res->n_expanded++;
        for (int a = 0; a < 12; ++a) {
            int lens[2];
            int st = aco_acmove(a, ns.states + cur * W, mrl, cyclical, child, lens);
            res->n_moves++;
            if (st != ACO_OK) {
                res->status = st;
                rc = st;
                goto out;
            }
            int nl = lens[0] + lens[1];
            last_len = nl;
            note_minlen(res, &min_length, nl);
            if (nl == 2) { /* greedy.py:91-100 */
                res->solved = 1;
                write_path(&ns, cur, a, nl, path, path_cap, res);
                goto out;
            }
            if (ns_find(&ns, child) < 0) { /* greedy.py:102-113 */
                int64_t k = ns_add(&ns, child, cur, a, nl, ns.depth[cur] + 1);
                heap_push(&hp, &ns, k);
            }
        }
        if (ns.n >= max_nodes) { /* greedy.py:115-119 */
            res->budget_hit = 1;
            break;
        }
    }
    /* greedy.py:121: (False, path + [(action, new_length)]) with the loop variables
     * of the last expanded node: action == 11, new_length of its 12th child */
    write_path(&ns, cur, 11, last_len, path, path_cap, res);
out:
    res->n_visited = ns.n;
    res->frontier_left = hp.n;
    dump_visited(&ns, visited_out, visited_cap);
    free(hp.h);
    ns_free(&ns);
    return rc;
}
