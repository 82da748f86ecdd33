// Poseidon permutation over Fr, as poseidon-rs 0.0.8 computes it behind the reference's
// POSEIDON.hash(..) (src/lib.rs:59, :333, :370, :401):  t = n_inputs + 1, R_F = 8, R_P(t),
// state = [0, in_0, ..],  per round:  add round constants -> x^5 on every lane (full rounds) or on
// lane 0 only (partial rounds) -> dense MDS mix;  output = state[0].
//
// The MDS row products are Montgomery dot products (fr_dot): t limb-products share ONE reduction.
#pragma once
#include "fr.cuh"

namespace bjj {

template <int T>
struct PoseidonTables;
#define BJJ_POSEIDON_TABLES(T_)                                                              \
    template <>                                                                              \
    struct PoseidonTables<T_> {                                                              \
        static BJJ_HD const uint32_t (*C())[8] { return BJJ_POSEIDON_C##T_; }                \
        static BJJ_HD const uint32_t (*M())[8] { return BJJ_POSEIDON_M##T_; }                \
        static constexpr int RP = BJJ_POSEIDON_RP_##T_;                                      \
    };
BJJ_POSEIDON_TABLES(2)
BJJ_POSEIDON_TABLES(3)
BJJ_POSEIDON_TABLES(4)
BJJ_POSEIDON_TABLES(5)
BJJ_POSEIDON_TABLES(6)
BJJ_POSEIDON_TABLES(7)
BJJ_POSEIDON_TABLES(8)
BJJ_POSEIDON_TABLES(9)

BJJ_HD void fr_pow5(Fr& x) {
    Fr x2, x4;
    fr_sqr(x2, x);
    fr_sqr(x4, x2);
    fr_mul(x, x4, x);
}

// state[] in Montgomery form, lazy domain; on return state[0] is the hash (Montgomery, lazy).
template <int T>
BJJ_HD void poseidon_permute(Fr* state) {
    const uint32_t(*C)[8] = PoseidonTables<T>::C();
    const uint32_t(*M)[8] = PoseidonTables<T>::M();
    constexpr int RP = PoseidonTables<T>::RP;
    constexpr int NR = 8 + RP;
#pragma unroll 1
    for (int r = 0; r < NR; r++) {
#pragma unroll
        for (int i = 0; i < T; i++) {
            Fr c = fr_const(C[r * T + i]);
            fr_add(state[i], state[i], c);
        }
        if (r < 4 || r >= 4 + RP) {
#pragma unroll
            for (int i = 0; i < T; i++) fr_pow5(state[i]);
        } else {
            fr_pow5(state[0]);
        }
        Fr ns[T];
#pragma unroll
        for (int i = 0; i < T; i++) {
            Fr row[T];
#pragma unroll
            for (int j = 0; j < T; j++) row[j] = fr_const(M[i * T + j]);
            fr_dot<T>(ns[i], row, state);
        }
#pragma unroll
        for (int i = 0; i < T; i++) state[i] = ns[i];
    }
}

}  // namespace bjj
