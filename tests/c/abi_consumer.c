/* A plain C99 consumer of include/asr.h: proves the boundary is a C ABI (no C++ or torch types in the
 * signatures), that the header compiles as C, and that the host-side checks answer without a GPU.
 * Built and run by tests/test_abi.py::test_plain_c_consumer. */
#include <stdio.h>
#include <string.h>

#include "asr.h"

int main(void) {
    size_t need = 0;
    AsrSolveParams p;
    memset(&p, 0, sizeof p);
    if (asr_version() != ASR_VERSION) { printf("version %d\n", asr_version()); return 1; }
    if (asr_solve_workspace_bytes(2, 100, 128, 128, 512, 512, 300, &need) != ASR_OK || need == 0) return 2;
    if (asr_solve_workspace_bytes(1, 4, 64, 64, 192, 192, 10, &need) != ASR_EUNSUPPORTED) return 3;   /* odd ratio */
    if (strstr(asr_last_error(), "feature_size") == NULL) return 4;
    if (asr_solve_workspace_bytes(2, 100, 128, 128, 512, 512, 300, NULL) != ASR_ENULL) return 5;
    /* null pointers are rejected before any CUDA call */
    if (asr_solve_batched(&p, 1, NULL, NULL, NULL, NULL, 1, 4, 16, 16, 64, 64, NULL, NULL, NULL, 0, NULL) != ASR_ENULL) return 6;
    printf("c abi ok: workspace for 2 x 100 copies = %zu bytes\n", need);
    return 0;
}
