/* oracle/shim/direct.h — stand-in for the MinGW <direct.h> the reference includes
 * (Algorithms/sequential/LZ4/LZ4.c:17); only _mkdir is used (S-LZ4:181-196). */
#ifndef ORACLE_SHIM_DIRECT_H
#define ORACLE_SHIM_DIRECT_H
#include <sys/stat.h>
#include <sys/types.h>
static inline int _mkdir(const char *p) { return mkdir(p, 0777); }
#endif
