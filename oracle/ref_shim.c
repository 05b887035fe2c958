/*
 * oracle/ref_shim.c - TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin accessors that get compiled together with the reference's own C
 * templates (src/{mc,itx,ipred,ipred_prepare}_tmpl.c, src/itx_1d.c,
 * src/tables.c, src/wedge.c, compiled in place from /root/reference by
 * oracle/Makefile) into oracle/_ref/libdav1d_ref.so.  The reference marks its
 * constant tables with hidden visibility (include/common/attributes.h:120),
 * so the parity tests reach them through this file.  Nothing here restates or
 * copies reference logic: it only hands out pointers.
 */
#include "config.h"
#include <stdint.h>
#include <string.h>
#include "common/attributes.h"
#include "src/tables.h"
#include "src/wedge.h"
#include "src/scan.h"

#define EXPORT __attribute__((visibility("default")))

struct tbl { const char *name; const void *ptr; size_t size; };

static const struct tbl *lookup(const char *name) {
    static const struct tbl tbls[] = {
        { "mc_subpel_filters",   dav1d_mc_subpel_filters,   sizeof(dav1d_mc_subpel_filters) },
        { "mc_warp_filter",      dav1d_mc_warp_filter,      sizeof(dav1d_mc_warp_filter) },
        { "resize_filter",       dav1d_resize_filter,       sizeof(dav1d_resize_filter) },
        { "sm_weights",          dav1d_sm_weights,          sizeof(dav1d_sm_weights) },
        { "dr_intra_derivative", dav1d_dr_intra_derivative, sizeof(dav1d_dr_intra_derivative) },
        { "filter_intra_taps",   dav1d_filter_intra_taps,   sizeof(dav1d_filter_intra_taps) },
        { "obmc_masks",          dav1d_obmc_masks,          sizeof(dav1d_obmc_masks) },
        { "txfm_dimensions",     dav1d_txfm_dimensions,     sizeof(dav1d_txfm_dimensions) },
        { "tx_type_class",       dav1d_tx_type_class,       sizeof(dav1d_tx_type_class) },
        { "block_dimensions",    dav1d_block_dimensions,    sizeof(dav1d_block_dimensions) },
        { "masks",               &dav1d_masks,              sizeof(dav1d_masks) },
    };
    for (size_t i = 0; i < sizeof(tbls) / sizeof(tbls[0]); i++)
        if (!strcmp(tbls[i].name, name)) return &tbls[i];
    return NULL;
}

EXPORT const void *oracle_ref_table(const char *name, size_t *size) {
    const struct tbl *t = lookup(name);
    if (!t) { if (size) *size = 0; return NULL; }
    if (size) *size = t->size;
    return t->ptr;
}

/* scan order for a (rect) tx size: uint16_t[], see src/scan.c:279-299 */
EXPORT const uint16_t *oracle_ref_scan(int tx) {
    return tx >= 0 && tx < N_RECT_TX_SIZES ? dav1d_scans[tx] : NULL;
}

EXPORT void oracle_ref_init_masks(void) {
    static int done;
    if (!done) { dav1d_init_ii_wedge_masks(); done = 1; }
}
