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

// state[] in Montgomery form, lazy domain; on return state[0] is the hash (Montgomery, lazy).
template <int T>
BJJ_HD void poseidon_permute(Fr* state) {
    const uint32_t(*FC)[8] = PoseidonTables<T>::FC();
    const uint32_t(*PRE)[8] = PoseidonTables<T>::PRE();
    const uint32_t(*PR)[8] = PoseidonTables<T>::PR();
    const uint32_t(*M)[8] = PoseidonTables<T>::M();
    const uint32_t(*P)[8] = PoseidonTables<T>::P();
    constexpr int RP = PoseidonTables<T>::RP;
#pragma unroll 1
    for (int r = 0; r < 4; r++) poseidon_full_round<T>(state, FC + r * T, r == 3 ? P : M);
#pragma unroll
    for (int i = 0; i < T; i++) {
        Fr c = fr_const(PRE[i]);
        fr_add(state[i], state[i], c);
    }
#pragma unroll 1
    for (int r = 0; r < RP; r++) {
        const uint32_t(*row)[8] = PR + r * (2 * T);
        Fr k = fr_const(row[0]);
        fr_add(state[0], state[0], k);
        fr_pow5(state[0]);
        Fr coef[T];
#pragma unroll
        for (int j = 0; j < T; j++) coef[j] = fr_const(row[1 + j]);      // a00, what[1..t-1]
        Fr n0;
        fr_dot<T>(n0, coef, state);
#pragma unroll
        for (int i = 1; i < T; i++) {
            Fr v = fr_const(row[T + i]), t;
            fr_mul(t, v, state[0]);
            fr_add(state[i], state[i], t);
        }
        state[0] = n0;
    }
#pragma unroll 1
    for (int r = 4; r < 8; r++) poseidon_full_round<T>(state, FC + r * T, M);
}

}  // namespace bjj
