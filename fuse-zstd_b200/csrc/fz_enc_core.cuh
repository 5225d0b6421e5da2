/*
 * fz_enc_core.cuh -- per-thread building blocks of the sm_100a zstd encoder that are also compiled for the host
 * by tests/emul (TEST ONLY there): FSE count normalisation, the FSE table description writer (the exact inverse of
 * read_ncount in fz_core.cuh, RFC 8878 4.1.1) and the encoding-table construction.
 */
#pragma once
#include "fz_core.cuh"

namespace fz {

// Accuracy log for a table coded from n_seq symbols with n_present distinct values (max_log: 9 LL / 8 OF / 9 ML).
FZ_HD int enc_table_log(uint32_t n_seq, int n_present, int max_log)
{
    int log = highbit(n_seq ? n_seq : 1) - 2;
    if (log < 5) log = 5;
    while ((1 << log) < 2 * n_present && log < max_log) log++;
    if (log > max_log) log = max_log;
    return log;
}

// count[0 .. n_sym) -> norm[] with sum exactly 1 << log and norm >= 1 for every present symbol.  Needs
// (number of present symbols) <= 1 << log and at least two present symbols.  Returns 0 or -1.
FZ_HD int enc_normalize(const uint32_t* count, int n_sym, uint32_t total, int log, int16_t* norm)
{
    const uint32_t T = 1u << log;
    int32_t sum = 0; int best = -1; uint32_t best_c = 0; int present = 0;
    for (int s = 0; s < n_sym; s++) {
        const uint32_t c = count[s];
        int32_t n = 0;
        if (c) { n = (int32_t)(((uint64_t)c << log) / total); if (n < 1) n = 1; present++; if (c > best_c) { best_c = c; best = s; } }
        norm[s] = (int16_t)n; sum += n;
    }
    if (present < 2 || (uint32_t)present > T) return -1;
    int32_t diff = (int32_t)T - sum;
    if (diff > 0) norm[best] = (int16_t)(norm[best] + diff);
    while (diff < 0) {                       // the minimum of 1 pushed the sum over: take it back from the largest entries
        int m = -1;
        for (int s = 0; s < n_sym; s++) if (norm[s] > 1 && (m < 0 || norm[s] > norm[m])) m = s;
        if (m < 0) return -1;
        const int32_t take = -diff < norm[m] - 1 ? -diff : norm[m] - 1;
        norm[m] = (int16_t)(norm[m] - take); diff += take;
    }
    return 0;
}

// FSE table description (RFC 8878 4.1.1) for norm[0 .. n_sym) (no "less than one" entries), accuracy `log`.
// Writes at most 2 + n_sym * 2 bytes at out; returns the byte count, or -1.
FZ_HD int enc_write_ncount(uint8_t* out, const int16_t* norm, int n_sym, int log)
{
    const int table_size = 1 << log;
    uint8_t* const start_out = out;
    uint32_t bits = (uint32_t)(log - 5); int nbits_held = 4;
    int remaining = table_size + 1, threshold = table_size, nb = log + 1;
    int sym = 0; bool prev0 = false;
    while (n_sym > 0 && norm[n_sym - 1] == 0) n_sym--;          // trailing absent symbols are implicit
    while (sym < n_sym && remaining > 1) {
        if (prev0) {
            int start = sym;
            while (sym < n_sym && norm[sym] == 0) sym++;
            if (sym == n_sym) break;
            while (sym >= start + 24) {
                start += 24; bits += 0xFFFFu << nbits_held;
                out[0] = (uint8_t)bits; out[1] = (uint8_t)(bits >> 8); out += 2; bits >>= 16;
            }
            while (sym >= start + 3) { start += 3; bits += 3u << nbits_held; nbits_held += 2; }
            bits += (uint32_t)(sym - start) << nbits_held; nbits_held += 2;
            if (nbits_held > 16) { out[0] = (uint8_t)bits; out[1] = (uint8_t)(bits >> 8); out += 2; bits >>= 16; nbits_held -= 16; }
        }
        {
            int count = norm[sym++];
            const int max = (2 * threshold - 1) - remaining;
            remaining -= count < 0 ? -count : count;
            count++;                                            // +1: the value 0 encodes "less than one"
            if (count >= threshold) count += max;
            bits += (uint32_t)count << nbits_held;
            nbits_held += nb;
            nbits_held -= (count < max);
            prev0 = (count == 1);
            if (remaining < 1) return -1;
            while (remaining < threshold) { nb--; threshold >>= 1; }
        }
        if (nbits_held > 16) { out[0] = (uint8_t)bits; out[1] = (uint8_t)(bits >> 8); out += 2; bits >>= 16; nbits_held -= 16; }
    }
    if (remaining != 1) return -1;
    out[0] = (uint8_t)bits; out[1] = (uint8_t)(bits >> 8);
    out += (nbits_held + 7) / 8;
    return (int)(out - start_out);
}

// FSE encoding table (next-state table + per-symbol transforms) from a normalised distribution; mirrors the
// decoder's spread.  state: 1 << log entries; dnb / dfs: n_sym entries; tmp: 1 << log bytes + (n_sym + 1) uint16.
FZ_HD void enc_build_ctable(uint16_t* state, uint32_t* dnb, int32_t* dfs, const int16_t* norm, int n_sym, int log, uint8_t* tmp_sym, uint16_t* cumul)
{
    const int size = 1 << log; int high = size - 1;
    cumul[0] = 0;
    for (int s = 0; s < n_sym; s++) {
        if (norm[s] == -1) { cumul[s + 1] = (uint16_t)(cumul[s] + 1); tmp_sym[high--] = (uint8_t)s; }
        else cumul[s + 1] = (uint16_t)(cumul[s] + norm[s]);
    }
    const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1; int pos = 0;
    for (int s = 0; s < n_sym; s++)
        for (int i = 0; i < norm[s]; i++) { tmp_sym[pos] = (uint8_t)s; do { pos = (pos + step) & mask; } while (pos > high); }
    for (int u = 0; u < size; u++) { const int s = tmp_sym[u]; state[cumul[s]++] = (uint16_t)(size + u); }
    int total = 0;
    for (int s = 0; s < n_sym; s++) {
        const int n = norm[s];
        if (n == 0) { dnb[s] = ((uint32_t)(log + 1) << 16) - (1u << log); dfs[s] = 0; }
        else if (n == -1 || n == 1) { dnb[s] = ((uint32_t)log << 16) - (1u << log); dfs[s] = total - 1; total++; }
        else {
            const uint32_t max_bits = (uint32_t)log - (uint32_t)highbit((uint32_t)n - 1);
            dnb[s] = (max_bits << 16) - ((uint32_t)n << max_bits); dfs[s] = total - n; total += n;
        }
    }
}

}  // namespace fz
