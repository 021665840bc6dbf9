/* Plain-C consumer of include/fdfd_b200.h: what a host language's FFI sees.  Compiled with `gcc -Iinclude` (no C++, no
 * CUDA headers) and linked against libfdfd_b200.so by tests/test_cabi_cpu.py.  Prints the layout of fdfd_desc /
 * fdfd_shape / fdfd_matparams_desc as `name offset` lines - the test compares them with the ctypes structures of
 * maxwellfdm.jl_b200/_lib.py and with the field order of the Julia struct in julia/FDFDB200.jl - then exercises the
 * entry points that need no GPU: version string, partition rule, halo plan, a host-only handle, and the loud failure
 * of every compute call on it. */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "fdfd_b200.h"

#define OFF(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))

int main(void) {
    OFF(fdfd_desc, N); OFF(fdfd_desc, isbloch); OFF(fdfd_desc, boundft_is_E); OFF(fdfd_desc, order_cmpfirst);
    OFF(fdfd_desc, field_type); OFF(fdfd_desc, device); OFF(fdfd_desc, rank); OFF(fdfd_desc, nranks);
    OFF(fdfd_desc, weighted_out_avg); OFF(fdfd_desc, kernel);
    printf("sizeof.fdfd_desc %zu\n", sizeof(fdfd_desc));
    OFF(fdfd_shape, kind); OFF(fdfd_shape, axis); OFF(fdfd_shape, pind); OFF(fdfd_shape, reserved); OFF(fdfd_shape, c);
    OFF(fdfd_shape, r);
    printf("sizeof.fdfd_shape %zu\n", sizeof(fdfd_shape));
    OFF(fdfd_matparams_desc, N); OFF(fdfd_matparams_desc, isbloch); OFF(fdfd_matparams_desc, boundft_is_E);
    OFF(fdfd_matparams_desc, field_type); OFF(fdfd_matparams_desc, field_ortho_shape); OFF(fdfd_matparams_desc, lprim);
    OFF(fdfd_matparams_desc, k0); OFF(fdfd_matparams_desc, k1); OFF(fdfd_matparams_desc, nshape);
    OFF(fdfd_matparams_desc, nparam); OFF(fdfd_matparams_desc, shapes); OFF(fdfd_matparams_desc, params);
    OFF(fdfd_matparams_desc, device);
    printf("sizeof.fdfd_matparams_desc %zu\n", sizeof(fdfd_matparams_desc));
    printf("sizeof.fdfd_c128 %zu\n", sizeof(fdfd_c128));

    if (!strstr(fdfd_version(), "sm_100a")) { printf("FAIL version %s\n", fdfd_version()); return 1; }

    int64_t k0 = -1, k1 = -1;
    if (fdfd_partition(768, 8, 3, &k0, &k1) != FDFD_OK || k0 != 288 || k1 != 384) { printf("FAIL partition\n"); return 1; }
    int32_t up = -2, dn = -2;
    if (fdfd_halo_plan(4, 0, 1, &up, &dn) != FDFD_OK || up != 1 || dn != 3) { printf("FAIL halo plan (wrap)\n"); return 1; }
    if (fdfd_halo_plan(4, 0, 0, &up, &dn) != FDFD_OK || up != 1 || dn != -1) { printf("FAIL halo plan\n"); return 1; }

    /* host-only handle: the debug export works, every GPU entry point refuses */
    fdfd_desc d;
    memset(&d, 0, sizeof d);
    d.N[0] = 3; d.N[1] = 2; d.N[2] = 2;
    d.isbloch[0] = d.isbloch[1] = d.isbloch[2] = 1;
    d.boundft_is_E[0] = d.boundft_is_E[1] = d.boundft_is_E[2] = 1;
    d.order_cmpfirst = 1;
    d.field_type = FDFD_FT_EE;
    d.device = -2;
    d.nranks = 1;
    fdfd_handle h = NULL;
    if (fdfd_create(&h, &d) != FDFD_OK || !h) { printf("FAIL create host-only: %s\n", fdfd_last_error(NULL)); return 1; }
    fdfd_c128 x[36], y[36];
    memset(x, 0, sizeof x);
    int rc = fdfd_apply(h, x, y, FDFD_HOST);
    if (rc == FDFD_OK) { printf("FAIL apply on a host-only handle returned OK\n"); return 1; }
    if (!fdfd_last_error(h) || !strlen(fdfd_last_error(h))) { printf("FAIL empty error message\n"); return 1; }
    if (fdfd_slab_range(h, &k0, &k1) != FDFD_OK || k0 != 0 || k1 != 2) { printf("FAIL slab range\n"); return 1; }
    fdfd_destroy(h);

    /* the single-call multi-GPU handle needs devices: without one it must fail loudly, not fall back */
    fdfd_multi m = NULL;
    d.device = -1;
    rc = fdfd_multi_create(&m, &d, 1, NULL);
    if (rc == FDFD_OK) {
        printf("multi: created on a GPU box (ngpu = %d)\n", fdfd_multi_ngpu(m));
        fdfd_multi_destroy(m);
    } else {
        const char *e = fdfd_multi_last_error(NULL);
        if (!e || !strlen(e)) { printf("FAIL multi_create failed without a message\n"); return 1; }
        printf("multi: refused without a GPU (%d)\n", rc);
    }
    printf("cabi_smoke ok\n");
    return 0;
}
