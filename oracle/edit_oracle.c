/* TEST INFRASTRUCTURE ONLY -- CPU checker for the edit-distance path.
 *
 * What it restates: CIRI_long/utils.py:153-159 `distance(x, y)`:
 *     Levenshtein.distance(x, y)               if len(x) <= 50 or len(y) <= 50
 *     edlib.align(x, y)['editDistance']        otherwise (edlib defaults: mode NW, unit costs, no
 *                                              extra equalities, k = -1)
 * Both third-party packages (python-Levenshtein, edlib; requirements of the reference, not vendored in
 * /root/reference and not installed in this image) compute the same quantity: the unit-cost global
 * (Needleman-Wunsch) edit distance between the two strings, symbols compared for exact equality.
 * PARITY PIN: the quantity is defined mathematically, so the checker is the textbook O(m*n) recurrence
 *     D[i][j] = min(D[i-1][j] + 1, D[i][j-1] + 1, D[i-1][j-1] + (x[i] != y[j]))
 * and tests/test_oracle.py pins it to published known answers and to metric properties.  Neither
 * edlib nor Levenshtein could be run here, which DESIGN.md states.
 */
#include <stdint.h>
#include <stdlib.h>

int32_t orc_edit_distance(const uint8_t* x, int32_t m, const uint8_t* y, int32_t n)
{
    if (m == 0) return n;
    if (n == 0) return m;
    int32_t* row = (int32_t*)malloc((size_t)(n + 1) * sizeof(int32_t));
    if (!row) return -1;
    for (int32_t j = 0; j <= n; ++j) row[j] = j;
    for (int32_t i = 1; i <= m; ++i) {
        int32_t diag = row[0];
        row[0] = i;
        for (int32_t j = 1; j <= n; ++j) {
            const int32_t up = row[j];
            int32_t best = diag + (x[i - 1] != y[j - 1]);
            if (up + 1 < best) best = up + 1;
            if (row[j - 1] + 1 < best) best = row[j - 1] + 1;
            diag = up;
            row[j] = best;
        }
    }
    const int32_t d = row[n];
    free(row);
    return d;
}

void orc_edit_distance_batch(int32_t n_pairs, const uint8_t* seqs, const int64_t* x_off, const int32_t* x_len,
                             const int64_t* y_off, const int32_t* y_len, int32_t* out)
{
    for (int32_t p = 0; p < n_pairs; ++p)
        out[p] = orc_edit_distance(seqs + x_off[p], x_len[p], seqs + y_off[p], y_len[p]);
}
