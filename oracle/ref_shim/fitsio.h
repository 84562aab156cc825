/* Shim for cfitsio's <fitsio.h>: the library is absent offline and FITS I/O is
 * out of scope. Only the handful of names the reference
 * mentions are declared; the harness (oracle/ref_harness.cu) defines the two
 * that end up referenced at link time. Test infrastructure only. */
#pragma once
#include <stdio.h>
typedef struct gvref_fitsfile { int unused; } fitsfile;
#ifdef __cplusplus
extern "C" {
#endif
int fits_read_key(fitsfile*, int, const char*, void*, char*, int*);
void fits_report_error(FILE*, int);
int fits_read_img(fitsfile*, int, long, long, void*, void*, int*, int*);
#ifdef __cplusplus
}
#endif
