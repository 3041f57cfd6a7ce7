// cf_tables.cpp -- see cf_tables.h
#include "cf_tables.h"

#include "joe_kuo_init.inc"

namespace cf {

int sobol_max_dim() { return JK_NDIM; }

// Joe & Kuo / Bratley & Fox recurrence.  For a primitive polynomial of degree s with interior
// coefficient bits a = (a_1 ... a_{s-1}) and odd initial values m_1..m_s (m_i < 2^i):
//   m_i = 2 a_1 m_{i-1} ^ 4 a_2 m_{i-2} ^ ... ^ 2^{s-1} a_{s-1} m_{i-s+1} ^ 2^s m_{i-s} ^ m_{i-s}
//   v_i = m_i << (32 - i)
const std::vector<uint32_t>& sobol_direction_table()
{
    static const std::vector<uint32_t> table = [] {
        const int nd = JK_NDIM;
        std::vector<uint32_t> t(size_t(32) * nd);
        for (int d = 0; d < nd; ++d) {
            const int s = JK_S[d];
            const unsigned a = JK_A[d];
            uint64_t mi[33];
            if (s == 0) {
                for (int i = 1; i <= 32; ++i) mi[i] = 1;
            } else {
                for (int i = 1; i <= s; ++i) mi[i] = JK_M[JK_OFF[d] + i - 1];
                for (int i = s + 1; i <= 32; ++i) {
                    uint64_t x = mi[i - s] ^ (mi[i - s] << s);
                    for (int k = 1; k < s; ++k)
                        if ((a >> (s - 1 - k)) & 1u) x ^= mi[i - k] << k;
                    mi[i] = x;
                }
            }
            for (int i = 1; i <= 32; ++i) t[size_t(i - 1) * nd + d] = uint32_t(mi[i] << (32 - i));
        }
        return t;
    }();
    return table;
}

namespace {
using u128 = unsigned __int128;

void matmul(const uint64_t a[9], const uint64_t b[9], uint64_t mod, uint64_t out[9])
{
    uint64_t t[9];
    for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) {
            u128 s = 0;
            for (int l = 0; l < 3; ++l) s += u128(a[3 * j + l]) * b[3 * l + k] % mod;
            t[3 * j + k] = uint64_t(s % mod);
        }
    for (int i = 0; i < 9; ++i) out[i] = t[i];
}

void matpow(const uint64_t a[9], uint64_t e, uint64_t mod, uint64_t out[9])
{
    uint64_t r[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, b[9];
    for (int i = 0; i < 9; ++i) b[i] = a[i];
    while (e) {
        if (e & 1) matmul(r, b, mod, r);
        matmul(b, b, mod, b);
        e >>= 1;
    }
    for (int i = 0; i < 9; ++i) out[i] = r[i];
}
}  // namespace

std::vector<uint64_t> mrg_jump_matrices(uint64_t stride)
{
    const uint64_t m1 = 4294967087ull, m2 = 4294944443ull;
    const uint64_t A[9] = {0, 1403580ull, m1 - 810728ull, 1, 0, 0, 0, 1, 0};
    const uint64_t B[9] = {527612ull, 0, m2 - 1370589ull, 1, 0, 0, 0, 1, 0};
    std::vector<uint64_t> out(size_t(32) * 2 * 9);
    uint64_t a[9], b[9];
    matpow(A, stride, m1, a);
    matpow(B, stride, m2, b);
    for (int k = 0; k < 32; ++k) {
        for (int i = 0; i < 9; ++i) { out[(size_t(k) * 2 + 0) * 9 + i] = a[i]; out[(size_t(k) * 2 + 1) * 9 + i] = b[i]; }
        matmul(a, a, m1, a);
        matmul(b, b, m2, b);
    }
    return out;
}

}  // namespace cf
