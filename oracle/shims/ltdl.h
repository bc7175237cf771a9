/* Build shim (test infrastructure, NOT product code).
 *
 * The reference's module.c / args_assembler.c include <ltdl.h> (GNU libltdl) to
 * dlopen filter plugins.  libltdl's development header is not installed in this
 * image and plugin loading is outside the hot path, so the
 * oracle build maps the handful of lt_* calls those files make onto plain
 * dlopen/dlsym.  Written from the libltdl API documentation, not copied from it.
 */
#ifndef ORACLE_SHIM_LTDL_H
#define ORACLE_SHIM_LTDL_H
#include <dlfcn.h>
#include <stddef.h>

typedef void *lt_dlhandle;
#define LT_PATHSEP_CHAR ':'

static inline int lt_dlinit(void) { return 0; }
static inline int lt_dlexit(void) { return 0; }
static inline int lt_dladdsearchdir(const char *dir) { (void) dir; return 0; }
static inline const char *lt_dlgetsearchpath(void) { return NULL; }
static inline int lt_dlforeachfile(const char *path, int (*func)(const char *filename, void *data), void *data) { (void) path; (void) func; (void) data; return 0; }
static inline int lt_dlsetsearchpath(const char *dir) { (void) dir; return 0; }
static inline lt_dlhandle lt_dlopenext(const char *name) { return dlopen(name, RTLD_NOW | RTLD_LOCAL); }
static inline lt_dlhandle lt_dlopen(const char *name) { return dlopen(name, RTLD_NOW | RTLD_LOCAL); }
static inline void *lt_dlsym(lt_dlhandle h, const char *sym) { return dlsym(h, sym); }
static inline int lt_dlclose(lt_dlhandle h) { return h ? dlclose(h) : 0; }
static inline const char *lt_dlerror(void) { return dlerror(); }

typedef struct { const char *filename; const char *name; int ref_count; } lt_dlinfo;
static inline const lt_dlinfo *lt_dlgetinfo(lt_dlhandle h) { (void) h; return NULL; }
#endif
