// scratch: trace active-set rounds on the host emulation
#include <cstdio>
#include <vector>
static int g_tr_nred[64], g_tr_nch[64], g_tr_rounds;
static int g_tr_act[64][64];
#define QR_TRACE_ROUND(round, nred, W) do { if (g_tr_rounds < 64) { g_tr_nred[g_tr_rounds] = nred; for (int f = 0; f < W.nf; ++f) g_tr_act[g_tr_rounds][f] = W.act[f]; ++g_tr_rounds; } } while (0)
#include "../tests/emul/emul.cpp"
extern "C" int trace_get(int* nred, int* acts, int* rounds) {
    for (int i = 0; i < 64; ++i) nred[i] = g_tr_nred[i];
    for (int i = 0; i < 64; ++i) for (int f = 0; f < 64; ++f) acts[i * 64 + f] = g_tr_act[i][f];
    *rounds = g_tr_rounds; g_tr_rounds = 0; return 0;
}
