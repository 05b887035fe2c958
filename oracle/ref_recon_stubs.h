/* Post-filter entry points that recon_tmpl.c's dav1d_filter_sbrow_*() drivers call (src/lf_apply.h,
 * src/cdef_apply.h, src/lr_apply.h).  The checker compiles recon_tmpl.c whole but only ever calls
 * dav1d_recon_b_intra / dav1d_backup_ipred_edge, so these are never reached. */
