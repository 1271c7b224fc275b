/* Plain-C consumer of include/anml_b200.h: proves the boundary is a C ABI
 * (no C++ or torch types in the signatures).  Built and run by
 * tests/test_cabi_symbols.py; needs no GPU. */
#include <stdio.h>
#include <string.h>

#include "anml_b200.h"

int main(void) {
    int count = -1;
    int rc = anml_device_count(&count);
    printf("version=%d device_count_rc=%d count=%d\n", anml_version(), rc, count);
    anml_design* d = NULL;
    /* argument validation happens before any CUDA call */
    if (anml_design_create(0, -5, 3, &d) != ANML_EINVAL) return 2;
    if (strlen(anml_last_error()) == 0) return 3;
    anml_model* m = NULL;
    if (anml_model_create(0, 10, ANML_FAMILY_GAUSSIAN_MEANVAR, 1, &m) != ANML_EINVAL) return 4;
    if (anml_design_destroy(NULL) != ANML_OK || anml_model_destroy(NULL) != ANML_OK) return 5;
    return anml_version() == ANML_B200_VERSION ? 0 : 1;
}
