// Poseidon permutation over Fr, as poseidon-rs 0.0.8 computes it behind the reference's
// POSEIDON.hash(..) (src/lib.rs:59, :333, :370, :401):  t = n_inputs + 1, R_F = 8, R_P(t),
// state = [0, in_0, ..],  per round:  add round constants -> x^5 on every lane (full rounds) or on
// lane 0 only (partial rounds) -> dense MDS mix;  output = state[0].
//
// The device evaluates an algebraically IDENTICAL schedule (exact field arithmetic => the same hash
// bits): the partial rounds use the sparse factorisation derived at build time by
// tools/gen_constants.py::poseidon_optimized -- per partial round one scalar constant, x^5 on lane 0,
// then   new0 = a00*x0 + <what, x[1..]>,   new_i = x_i + v_i*x0   (2t-1 products instead of t^2).
// Every row product is a Montgomery dot product (fr_dot): the limb products share ONE reduction.
#pragma once
#include "fr.cuh"

namespace bjj {

template <int T>
struct PoseidonTables;
#define BJJ_POSEIDON_TABLES(T_)                                                              \
    template <>                                                                              \
    struct PoseidonTables<T_> {                                                              \
        static BJJ_HD const uint32_t (*FC())[8] { return BJJ_POSEIDON_FC##T_; }              \
        static BJJ_HD const uint32_t (*PRE())[8] { return BJJ_POSEIDON_PRE##T_; }            \
        static BJJ_HD const uint32_t (*PR())[8] { return BJJ_POSEIDON_PR##T_; }              \
        static BJJ_HD const uint32_t (*M())[8] { return BJJ_POSEIDON_M##T_; }                \
        static BJJ_HD const uint32_t (*P())[8] { return BJJ_POSEIDON_P##T_; }                \
        static constexpr int RP = BJJ_POSEIDON_RP_##T_;                                      \
        static constexpr int GROUP = BJJ_POSEIDON_GROUP_##T_;                                \
    };
BJJ_POSEIDON_TABLES(2)
BJJ_POSEIDON_TABLES(3)
BJJ_POSEIDON_TABLES(4)
BJJ_POSEIDON_TABLES(5)
BJJ_POSEIDON_TABLES(6)
BJJ_POSEIDON_TABLES(7)

BJJ_HD void fr_pow5(Fr& x) {
    Fr x2, x4;
    fr_sqr(x2, x);
    fr_sqr(x4, x2);
    fr_mul(x, x4, x);
}

// one full round: add constants, x^5 on every lane, dense mix with `mat`
template <int T>
BJJ_HD void poseidon_full_round(Fr* state, const uint32_t (*rc)[8], const uint32_t (*mat)[8]) {
#pragma unroll
    for (int i = 0; i < T; i++) {
        Fr c = fr_const(rc[i]);
        fr_add(state[i], state[i], c);
        fr_pow5(state[i]);
    }
    Fr ns[T];
#pragma unroll
    for (int i = 0; i < T; i++) {
        Fr row[T];
#pragma unroll
        for (int j = 0; j < T; j++) row[j] = fr_const(mat[i * T + j]);
        fr_dot<T>(ns[i], row, state);
    }
#pragma unroll
    for (int i = 0; i < T; i++) state[i] = ns[i];
}

// K consecutive partial rounds (tools/gen_constants.py::poseidon_groups): lanes 1..T-1 only accumulate v_i * x0
// between partial rounds, so inside a group lane 0 is one dot product over the lanes as they were at the START of the
// group and the x0 of the group's earlier rounds (T + j terms, one reduction), and lanes 1..T-1 are brought up to date
// once per group by a K-term dot product each: 14.7 products-or-reductions of 64 MACs per round instead of 17 (T = 6).
// round J of a group: x0 = (lane 0 + k)^5, then lane 0 = <(a00, c_(J,J-1), .., c_(J,0), what), (x0_J, .., x0_0, S_1, ..)>
template <int T, int K, int J>
BJJ_HD void poseidon_group_round(Fr* state, Fr* x, const uint32_t (*g)[8]) {
    Fr k = fr_const(g[0]);
    fr_add(state[0], state[0], k);
    fr_pow5(state[0]);
    x[J] = state[0];
    Fr coef[T + J], ops[T + J];
#pragma unroll
    for (int m = 0; m <= J; m++) {
        coef[m] = fr_const(g[1 + m]);
        ops[m] = x[J - m];
    }
#pragma unroll
    for (int i = 1; i < T; i++) {
        coef[J + i] = fr_const(g[1 + J + i]);
        ops[J + i] = state[i];
    }
    fr_dot<T + J>(state[0], coef, ops);
    if constexpr (J + 1 < K) poseidon_group_round<T, K, J + 1>(state, x, g + T + 1 + J);
}

template <int T, int K>
BJJ_HD void poseidon_partial_group(Fr* state, const uint32_t (*g)[8]) {
    Fr x[K];
    poseidon_group_round<T, K, 0>(state, x, g);
    g += K * (T + 1) + K * (K - 1) / 2;
#pragma unroll
    for (int i = 1; i < T; i++) {
        Fr t;
        if constexpr (K == 1) {
            Fr v = fr_const(g[0]);
            fr_mul(t, v, x[0]);
        } else {
            Fr coef[K];
#pragma unroll
            for (int m = 0; m < K; m++) coef[m] = fr_const(g[m]);
            fr_dot<K>(t, coef, x);
        }
        fr_add(state[i], state[i], t);
        g += K;
    }
}

// state[] in Montgomery form, lazy domain; on return state[0] is the hash (Montgomery, lazy).
template <int T>
BJJ_HD void poseidon_permute(Fr* state) {
    const uint32_t(*FC)[8] = PoseidonTables<T>::FC();
    const uint32_t(*PRE)[8] = PoseidonTables<T>::PRE();
    const uint32_t(*PR)[8] = PoseidonTables<T>::PR();
    const uint32_t(*M)[8] = PoseidonTables<T>::M();
    const uint32_t(*P)[8] = PoseidonTables<T>::P();
    constexpr int RP = PoseidonTables<T>::RP;
    constexpr int G = PoseidonTables<T>::GROUP;       // partial rounds per group: 3 for t >= 5, else 1 (gen_constants.py)
    static_assert(G == 1 || G == 3, "poseidon_partial_group handles groups of 1..3");
#pragma unroll 1
    for (int r = 0; r < 4; r++) poseidon_full_round<T>(state, FC + r * T, r == 3 ? P : M);
#pragma unroll
    for (int i = 0; i < T; i++) {
        Fr c = fr_const(PRE[i]);
        fr_add(state[i], state[i], c);
    }
#pragma unroll 1
    for (int r = 0; r < RP / G; r++) poseidon_partial_group<T, G>(state, PR + r * (2 * G * T + G * (G - 1) / 2));
    if constexpr (RP % G == 2) poseidon_partial_group<T, 2>(state, PR + (RP / G) * (2 * G * T + G * (G - 1) / 2));
    if constexpr (RP % G == 1) poseidon_partial_group<T, 1>(state, PR + (RP / G) * (2 * G * T + G * (G - 1) / 2));
#pragma unroll 1
    for (int r = 4; r < 7; r++) poseidon_full_round<T>(state, FC + r * T, M);
    // last round: only lane 0 leaves the permutation, so only row 0 of the mix is computed
    {
        const uint32_t(*rc)[8] = FC + 7 * T;
#pragma unroll
        for (int i = 0; i < T; i++) {
            Fr c = fr_const(rc[i]);
            fr_add(state[i], state[i], c);
            fr_pow5(state[i]);
        }
        Fr row[T], h;
#pragma unroll
        for (int j = 0; j < T; j++) row[j] = fr_const(M[j]);
        fr_dot<T>(h, row, state);
        state[0] = h;
    }
}

}  // namespace bjj
