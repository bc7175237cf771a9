/* Build shim (test infrastructure, NOT product code).
 *
 * The reference's writer.c / fileio.c include <bzlib.h>; the bzip2 runtime
 * (libbz2.so.1.0) is present in this image but its development header is not.
 * Compressed I/O is outside the hot path, so the oracle build only
 * needs the prototypes of the high-level stdio-style calls those two files make.
 * Prototypes restated from the bzip2 manual (section 3.4, "High(er) level
 * library functions").
 */
#ifndef ORACLE_SHIM_BZLIB_H
#define ORACLE_SHIM_BZLIB_H
#include <stdio.h>
typedef void BZFILE;
#define BZ_OK 0
#define BZ_STREAM_END 4
BZFILE *BZ2_bzopen(const char *path, const char *mode);
BZFILE *BZ2_bzdopen(int fd, const char *mode);
int BZ2_bzread(BZFILE *b, void *buf, int len);
int BZ2_bzwrite(BZFILE *b, void *buf, int len);
int BZ2_bzflush(BZFILE *b);
void BZ2_bzclose(BZFILE *b);
const char *BZ2_bzerror(BZFILE *b, int *errnum);
#endif
