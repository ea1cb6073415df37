/* util_quda.h -- drop-in stand-in for upstream QUDA's logging header: printfQuda (rank 0 only, silenced by QUDA_SILENT), warningQuda,
 * errorQuda (prints file / line and aborts the job), as the reference drivers and the plug-in use them (SURVEY.md 8b "Errors"). */
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include "quda.h"
#ifdef __cplusplus
extern "C" {
#endif
QudaVerbosity getVerbosity(void);
void qkxtm_error_at(const char *file, int line, const char *func, const char *fmt, ...);   /* aborts (or calls the installed handler) */
#ifdef __cplusplus
}
#endif
#define printfQuda(...) do { if (getVerbosity() > QUDA_SILENT && comm_rank() == 0) { printf(__VA_ARGS__); fflush(stdout); } } while (0)
#define warningQuda(...) do { if (getVerbosity() > QUDA_SILENT && comm_rank() == 0) { fprintf(stderr, "WARNING: "); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)
#define errorQuda(...) qkxtm_error_at(__FILE__, __LINE__, __func__, __VA_ARGS__)
