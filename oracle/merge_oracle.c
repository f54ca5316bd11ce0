/*
 * merge_oracle.c — CPU ORACLE of the in-order merge of resolved rows. TEST INFRASTRUCTURE ONLY (same rules as tplx_oracle.c).
 *
 * Literal restatement of ResolveTask::executeInOrder + processExceptionRow + emitNormalRows
 * (/root/reference/tuplex/core/src/physical/ResolveTask.cc:878-1258, :396-470, :313-375) for runtime exceptions of one task:
 *   for every exception record in row-number order:
 *       _currentRowNumber = the record's row number
 *       emitNormalRows():  while (_rowNumber != _currentRowNumber) { write the next normal row; _rowNumber++ }     (:324-348)
 *       resolve the row; if it resolved, write the resolved row                                                     (:252-261)
 *       _rowNumber++   — the exception's slot is consumed whether or not it resolved                                (:971)
 *   afterwards the remaining normal rows are written                                                                (:1180-1258)
 * Output: for every row of the merged stream where it comes from — a normal row index, or ~m for the m-th RESOLVED exception
 * (m counts resolved exceptions in row-number order). Checker for K9 (tuplex_b200/csrc/merge.cuh), which uses a closed form.
 */
#include <stdint.h>

/* exc_row_nos: ascending row numbers of ALL exception records of the task (relative to the task's first row number);
 * resolved[k] != 0: exception k produced a row. Returns the number of merged rows written to src (capacity n_norm + n_exc). */
uint64_t tplx_oracle_merge(uint64_t n_norm, const int64_t *exc_row_nos, const uint8_t *resolved, uint64_t n_exc, int64_t *src) {
    uint64_t out = 0, normal = 0, m = 0;
    int64_t row_number = 0;
    for (uint64_t k = 0; k < n_exc; ++k) {
        const int64_t current = exc_row_nos[k];
        while (row_number != current && normal < n_norm) { /* emitNormalRows */
            src[out++] = (int64_t)normal++;
            row_number++;
        }
        if (resolved[k]) src[out++] = ~(int64_t)m++;
        row_number++;
    }
    while (normal < n_norm) src[out++] = (int64_t)normal++;
    return out;
}
