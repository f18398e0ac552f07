/*
 * TEST / MEASUREMENT INFRASTRUCTURE (oracle) -- not part of the product path.
 *
 * Op-counting build of oracle/mek_oracle.c ("exact tallies from an op-counting build of the
 * structured oracle").  The C source is compiled unchanged as C++ with `double` replaced by a scalar class that
 * performs the same IEEE-754 binary64 operation and counts it, attributed to the block the source marked last with
 * OPC_BLOCK(id).  Counting rule: add/sub = 1, mul = 1, fma = 2, div = 1, sqrt = 1 flop; negation,
 * fabs, floor, comparisons, max/select and int<->double conversions are free (comparisons are tallied separately).
 * Constant expressions between literals (1.0 / 23.0) are folded by the compiler in both builds and not counted.
 *
 * Build (oracle/Makefile): g++ -O1 -ffp-contract=off -shared -fPIC opcount.cpp -o _build/libmek_oracle_count.so
 * Single-threaded (the OpenMP pragmas of the source are ignored: counters are plain globals).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { OPC_NBLOCK = 10, OPC_NKIND = 6 };            /* kinds: add, mul, fma, div, sqrt, cmp */
static unsigned long long g_cnt[OPC_NBLOCK][OPC_NKIND];
static int g_blk = OPC_NBLOCK - 1;                  /* last block = "outside any marked block" (init, run constants) */
#define OPC_BLOCK(id) (g_blk = (id))
#define OPC_TICK(kind) (++g_cnt[g_blk][kind])

struct cnt {
  double v;
  cnt() = default;
  cnt(double x) : v(x) {}
  template <class T> explicit operator T() const { return (T)v; }
};
static inline cnt operator+(cnt a, cnt b) { OPC_TICK(0); return cnt(a.v + b.v); }
static inline cnt operator-(cnt a, cnt b) { OPC_TICK(0); return cnt(a.v - b.v); }
static inline cnt operator*(cnt a, cnt b) { OPC_TICK(1); return cnt(a.v * b.v); }
static inline cnt operator/(cnt a, cnt b) { OPC_TICK(3); return cnt(a.v / b.v); }
static inline cnt operator-(cnt a) { return cnt(-a.v); }
static inline cnt &operator+=(cnt &a, cnt b) { OPC_TICK(0); a.v += b.v; return a; }
static inline cnt &operator-=(cnt &a, cnt b) { OPC_TICK(0); a.v -= b.v; return a; }
static inline cnt &operator*=(cnt &a, cnt b) { OPC_TICK(1); a.v *= b.v; return a; }
static inline bool operator<(cnt a, cnt b) { OPC_TICK(5); return a.v < b.v; }
static inline bool operator>(cnt a, cnt b) { OPC_TICK(5); return a.v > b.v; }
static inline bool operator<=(cnt a, cnt b) { OPC_TICK(5); return a.v <= b.v; }
static inline bool operator>=(cnt a, cnt b) { OPC_TICK(5); return a.v >= b.v; }
static inline bool operator==(cnt a, cnt b) { OPC_TICK(5); return a.v == b.v; }
static inline bool operator!=(cnt a, cnt b) { OPC_TICK(5); return a.v != b.v; }
static inline cnt fma(cnt a, cnt b, cnt c) { OPC_TICK(2); return cnt(::fma(a.v, b.v, c.v)); }
static inline cnt sqrt(cnt a) { OPC_TICK(4); return cnt(::sqrt(a.v)); }
static inline cnt fabs(cnt a) { return cnt(::fabs(a.v)); }
static inline cnt floor(cnt a) { return cnt(::floor(a.v)); }
static inline cnt fmax(cnt a, cnt b) { OPC_TICK(5); return cnt(::fmax(a.v, b.v)); }
static inline cnt fmin(cnt a, cnt b) { OPC_TICK(5); return cnt(::fmin(a.v, b.v)); }

extern "C" {
void mko_opc_reset(void) { memset(g_cnt, 0, sizeof g_cnt); g_blk = OPC_NBLOCK - 1; }
/* out[block][kind], unsigned long long */
void mko_opc_get(unsigned long long *out) { memcpy(out, g_cnt, sizeof g_cnt); }
int mko_opc_dims(int *nblock, int *nkind) { *nblock = OPC_NBLOCK; *nkind = OPC_NKIND; return 0; }

#define double cnt
#include "mek_oracle.c"
#undef double
}
