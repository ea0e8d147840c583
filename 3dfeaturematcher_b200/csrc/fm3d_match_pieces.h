// Work distribution of the persistent tensor-core matchers (match_tc_persistent_kernel, match_sp_persistent_kernel):
// the (query tile, train tile) pairs, query-tile-major, cut into G equal contiguous ranges -- one per CTA.  A range is a
// sequence of "pieces": (query tile, first train tile, number of train tiles, list slot).  The pieces of one query tile
// belong to consecutive CTAs and take list slots 0, 1, ...  Plain C++ so that the host (launch plan) and a CPU test
// (tests/cpp/pieces_harness.cpp) walk exactly what the kernels walk.
#pragma once

#if defined(__CUDACC__)
#define FM3D_PIECES_HD __host__ __device__
#else
#define FM3D_PIECES_HD
#endif

// the CTA whose range [c * total / G, (c + 1) * total / G) holds flattened tile y
FM3D_PIECES_HD inline int tcp_cta_of(long long y, long long total, int G) { return (int)(((y + 1) * G - 1) / total); }

struct TcpPieces {                // the pieces of one CTA, in order; every role of the kernel walks its own copy
    int G, cta, nt_tiles;
    long long total, T, T1;
    FM3D_PIECES_HD TcpPieces(int q_tiles, int nt_tiles_, int G_, int cta_) {
        G = G_; cta = cta_; nt_tiles = nt_tiles_;
        total = (long long)q_tiles * nt_tiles;             // the host launches G <= total CTAs: no range is empty
        T = (long long)cta * total / G;
        T1 = (long long)(cta + 1) * total / G;
    }
    FM3D_PIECES_HD bool next(int& qtile, int& tile_lo, int& ntiles, int& slot) {
        if (T >= T1) return false;
        qtile = (int)(T / nt_tiles);
        tile_lo = (int)(T - (long long)qtile * nt_tiles);
        const long long rest_q = (long long)(nt_tiles - tile_lo), rest_r = T1 - T;
        const long long n = rest_q < rest_r ? rest_q : rest_r;
        ntiles = (int)n;
        slot = cta - tcp_cta_of((long long)qtile * nt_tiles, total, G);
        T += n;
        return true;
    }
};

// The launch plan: CTAs (every SM, or as many as get `min_tiles` tiles each) and the most CTAs a query tile's train
// sequence is cut over (= list slots per query row and epilogue half).
inline void tcp_plan(int q_tiles, int nt_tiles, int sms, int min_tiles, int* G_out, int* pieces_out) {
    const long long total = (long long)q_tiles * nt_tiles;
    long long G = total / (min_tiles > 0 ? min_tiles : 1);
    G = G < 1 ? 1 : (G > sms ? sms : G);
    int pieces = 1;
    for (int qt = 0; qt < q_tiles; qt++) {
        const int c0 = tcp_cta_of((long long)qt * nt_tiles, total, (int)G), c1 = tcp_cta_of((long long)(qt + 1) * nt_tiles - 1, total, (int)G);
        if (c1 - c0 + 1 > pieces) pieces = c1 - c0 + 1;
    }
    *G_out = (int)G;
    *pieces_out = pieces;
}
