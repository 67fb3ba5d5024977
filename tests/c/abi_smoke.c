/* The boundary is a C ABI: this file is compiled as C (gcc -std=c99, not C++) against include/ocm_b200.h and linked
 * with libocm_b200.so.  It only touches host-side entry points, so it runs without a GPU. */
#include <stdio.h>
#include <string.h>

#include "ocm_b200.h"

int main(void) {
    const double mass[4] = {12.0, 12.0, 14.0, 16.0};
    const double box[3] = {5.0, 5.0, 5.0};
    const int32_t ij[4] = {0, 1, 2, 3};
    const double r0[2] = {0.38, 0.38}, k[2] = {1000.0, 1000.0};
    const int32_t bad[2] = {0, 7};
    ocm_system* sys = NULL;
    if (!strstr(ocm_version(), "ocm-b200")) return 1;
    if (ocm_device_count() < 0) return 2;
    if (ocm_system_create(4, mass, box, &sys) != 0 || sys == NULL) return 3;
    if (ocm_system_num_particles(sys) != 4) return 4;
    if (ocm_system_add_harmonic_bond(sys, 0, 1, 2, ij, r0, k) != 0) return 5;
    if (ocm_system_add_harmonic_bond(sys, 0, 0, 1, bad, r0, k) == 0) return 6; /* index out of range must fail */
    if (strlen(ocm_last_error()) == 0) return 7;
    ocm_system_destroy(sys);
    printf("abi ok: %s, %d device(s)\n", ocm_version(), ocm_device_count());
    return 0;
}
