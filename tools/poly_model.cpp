// tools/poly_model.cpp -- host model of the polylines kernel's per-column logic (development aid, not shipped).
// Builds the sorted tables of one row the simple way (std::stable_sort), then runs the SAME classification, float32
// fast path and exact path the CUDA kernel runs (csrc/cs_poly_core.cuh).  tools/poly_model_check.py compares the exact
// path with the oracle and the certified fast path with the exact path.
//   g++ -O2 -ffp-contract=off -shared -fPIC -o /tmp/libpoly_model.so tools/poly_model.cpp
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "../comfystereo_b200/csrc/cs_poly_core.cuh"

using namespace cs::poly;

template <bool SHARP>
static int run_row(const uint8_t* img, const float* nd, int w, double div_px, double sep_px, double expo,
                   uint8_t* out_exact, uint8_t* out_fast, uint8_t* fast_ok, int* stats) {
    const int npts = (SHARP ? 2 * w : w) + 2, nsg = npts - 1;
    if (npts > 65535) return -1;
    std::vector<float> X(npts), SX(npts), Q(w + 2);
    std::vector<uint32_t> ER(npts), IMG(npts);
    std::vector<uint16_t> SID(npts), WIN(npts), START(w + 4), RNK(npts);
    X[0] = (float)(-1.0 * w);
    X[npts - 1] = (float)(2.0 * w);
    Q[0] = 0.0f; Q[w + 1] = 0.0f;
    for (int col = 0; col < w; ++col) {
        const float d = nd[col];
        const double an = (double)fabsf(d);
        const double p = (expo == 2.0) ? an * an : (expo == 1.0 ? an : pow(an, expo));
        const double sp = (d >= 0.0f) ? p : -p;
        const double cd = sp * div_px;
        double cx = ((double)col + 0.5) + cd;
        cx = cx + sep_px;
        Q[col + 1] = (float)fabs(cd);
        if (SHARP) { X[1 + 2 * col] = (float)(cx - 0.45); X[2 + 2 * col] = (float)(cx + 0.45); }
        else X[1 + col] = (float)cx;
        const uint32_t px = (uint32_t)img[3 * col] | ((uint32_t)img[3 * col + 1] << 8) | ((uint32_t)img[3 * col + 2] << 16);
        if (SHARP) { IMG[1 + 2 * col] = px; IMG[2 + 2 * col] = px; }
        else IMG[1 + col] = px;
    }
    IMG[0] = IMG[1]; IMG[npts - 1] = IMG[npts - 2];
    std::vector<int> order(npts);
    for (int i = 0; i < npts; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return X[a] < X[b]; });
    for (int k = 0; k < npts; ++k) { RNK[order[k]] = (uint16_t)k; SX[k] = X[order[k]]; SID[k] = (uint16_t)order[k]; }
    uint32_t reach = 0;
    for (int k = 0; k < npts; ++k) {
        const int i = order[k];
        const uint32_t end = (i < nsg) ? RNK[i + 1] : 0u;
        reach = std::max(reach, end);
        ER[k] = end | (reach << 16);
    }
    const int tw = w;
    int bprev = -1;
    for (int k = 0; k < npts; ++k) {
        const int b = std::min(std::max((int)floorf(SX[k]) + 1, 0), tw + 1);
        if (b > bprev) { for (int q = bprev + 1; q <= b; ++q) START[q] = (uint16_t)k; bprev = b; }
    }
    START[tw + 2] = (uint16_t)npts;
    Tab t;
    t.X = X.data(); t.SX = SX.data(); t.ER = ER.data(); t.SID = SID.data(); t.WSP = WIN.data(); t.Q = Q.data();
    t.IMGP = IMG.data(); t.START = START.data(); t.w = w; t.npts = npts; t.nsg = nsg; t.t0 = 0;
    int nhard = 0, ncode[4] = {0, 0, 0, 0};
    for (int k = 0; k < nsg; ++k) {
        const int rprev = k ? (int)(ER[k - 1] >> 16) : 0;
        uint32_t code;
        if (rprev <= k) code = ((int)(ER[k] & 0xFFFFu) > k) ? (uint32_t)SID[k] : ((uint32_t)SID[k] | kUnresolved);
        else { code = classify_interval<SHARP>(t, k); ++nhard; }
        WIN[k] = (uint16_t)code;
        ++ncode[(code & kUnresolved) ? 3 : 1];
    }
    WIN[nsg] = 0;
    int gave_up = 0, nbad = 0;
    for (int col = 0; col < w; ++col) {
        uint32_t pe = 0, pf = 0;
        pe = exact_column<SHARP>(t, col);
        if (pe & kGaveUp) gave_up = 1;
        const bool ok = fast_column<SHARP>(t, col, &pf);
        for (int ch = 0; ch < 3; ++ch) {
            out_exact[3 * col + ch] = (uint8_t)(pe >> (8 * ch));
            out_fast[3 * col + ch] = ok ? (uint8_t)(pf >> (8 * ch)) : 0;
        }
        fast_ok[col] = ok ? 1 : 0;
        nbad += ok ? 0 : 1;
    }
    if (stats) { stats[0] += nhard; stats[1] += ncode[0]; stats[2] += ncode[1]; stats[3] += ncode[2]; stats[4] += ncode[3];
                 stats[5] += nbad; stats[6] += gave_up; stats[7] += nsg; }
    return gave_up;
}

extern "C" int poly_model_row(const uint8_t* img, const float* nd, int w, double div_px, double sep_px, double expo,
                              int sharp, uint8_t* out_exact, uint8_t* out_fast, uint8_t* fast_ok, int* stats) {
    return sharp ? run_row<true>(img, nd, w, div_px, sep_px, expo, out_exact, out_fast, fast_ok, stats)
                 : run_row<false>(img, nd, w, div_px, sep_px, expo, out_exact, out_fast, fast_ok, stats);
}
