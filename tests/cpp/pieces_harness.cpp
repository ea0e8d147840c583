// Test-only host harness: walks the work distribution of the persistent matchers (csrc/fm3d_match_pieces.h) exactly as the
// kernels do and reports what tests/test_abi_and_host.py checks: every (query tile, train tile) pair exactly once, list
// slots of a query tile consecutive from 0 and below the planned count, ranges balanced to one tile.  Not part of libfm3d.
#include <vector>
#include <cstdint>
#include "fm3d_match_pieces.h"

// returns 0 if everything holds, else a code saying what failed; out = {G, pieces, min tiles per CTA, max tiles per CTA}
extern "C" int pieces_check(int q_tiles, int nt_tiles, int sms, int min_tiles, int* out) {
    int G = 0, pieces = 0;
    tcp_plan(q_tiles, nt_tiles, sms, min_tiles, &G, &pieces);
    const long long total = (long long)q_tiles * nt_tiles;
    if (G < 1 || G > sms || G > total) return 1;
    std::vector<uint8_t> seen((size_t)total, 0);
    std::vector<int> next_slot(q_tiles, 0), closed(q_tiles, 0);
    long long wmin = total, wmax = 0;
    for (int cta = 0; cta < G; cta++) {
        TcpPieces it(q_tiles, nt_tiles, G, cta);
        int q, lo, n, slot;
        long long w = 0;
        int last_q = -1;
        while (it.next(q, lo, n, slot)) {
            if (q < 0 || q >= q_tiles || lo < 0 || n < 1 || lo + n > nt_tiles) return 2;
            if (q <= last_q) return 3;                         // a CTA meets a query tile once, in ascending order
            last_q = q;
            if (slot != next_slot[q] || slot >= pieces) return 4;   // consecutive CTAs -> consecutive slots
            next_slot[q]++;
            if (closed[q]) return 5;
            if (lo + n == nt_tiles) closed[q] = 1;             // the piece that marks the slots above as empty
            for (int t = lo; t < lo + n; t++) {
                uint8_t& s = seen[(size_t)q * nt_tiles + t];
                if (s) return 6;
                s = 1;
            }
            w += n;
        }
        if (w < 1) return 7;
        wmin = w < wmin ? w : wmin; wmax = w > wmax ? w : wmax;
    }
    for (long long i = 0; i < total; i++) if (!seen[(size_t)i]) return 8;
    for (int q = 0; q < q_tiles; q++) if (!closed[q]) return 9;
    if (wmax - wmin > 1) return 10;
    out[0] = G; out[1] = pieces; out[2] = (int)wmin; out[3] = (int)wmax;
    return 0;
}
