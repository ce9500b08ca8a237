/* oracle/shim/windows.h — TEST INFRASTRUCTURE.  The handful of Win32 names the reference's parallel builds use
 * (Algorithms/parallel/LZ4/LZ4.c:16, Algorithms/parallel/JPEG/JPEG.c:8), mapped onto pthreads so that those sources
 * compile unmodified with gcc on Linux.  Nothing in the product includes this file. */
#ifndef LJB_SHIM_WINDOWS_H
#define LJB_SHIM_WINDOWS_H
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

typedef unsigned long DWORD;
typedef void *LPVOID;
typedef int BOOL;
#define WINAPI
#define TRUE 1
#define FALSE 0
#define INFINITE 0xFFFFFFFFu
#define WAIT_OBJECT_0 0

typedef struct ljb_shim_thread {
    pthread_t th;
    DWORD (*fn)(LPVOID);
    LPVOID arg;
} *HANDLE;
typedef DWORD (*LPTHREAD_START_ROUTINE)(LPVOID);

static void *ljb_shim_thread_main(void *p)
{
    HANDLE h = (HANDLE)p;
    h->fn(h->arg);
    return NULL;
}
static inline HANDLE CreateThread(void *attr, size_t stack, LPTHREAD_START_ROUTINE fn, LPVOID arg, DWORD flags, DWORD *id)
{
    (void)attr; (void)stack; (void)flags;
    HANDLE h = (HANDLE)malloc(sizeof *h);
    if (!h) return NULL;
    h->fn = fn;
    h->arg = arg;
    if (pthread_create(&h->th, NULL, ljb_shim_thread_main, h) != 0) {
        free(h);
        return NULL;
    }
    if (id) *id = 0;
    return h;
}
static inline DWORD WaitForSingleObject(HANDLE h, DWORD ms)
{
    (void)ms;
    if (h) pthread_join(h->th, NULL);
    return WAIT_OBJECT_0;
}
static inline DWORD WaitForMultipleObjects(DWORD n, HANDLE *hs, BOOL all, DWORD ms)
{
    (void)all; (void)ms;
    for (DWORD i = 0; i < n; ++i) WaitForSingleObject(hs[i], INFINITE); /* (no 64-handle limit here) */
    return WAIT_OBJECT_0;
}
static inline BOOL CloseHandle(HANDLE h) { free(h); return TRUE; }

typedef pthread_mutex_t CRITICAL_SECTION;
static inline void InitializeCriticalSection(CRITICAL_SECTION *c) { pthread_mutex_init(c, NULL); }
static inline void DeleteCriticalSection(CRITICAL_SECTION *c) { pthread_mutex_destroy(c); }
static inline void EnterCriticalSection(CRITICAL_SECTION *c) { pthread_mutex_lock(c); }
static inline void LeaveCriticalSection(CRITICAL_SECTION *c) { pthread_mutex_unlock(c); }

typedef struct { DWORD dwNumberOfProcessors; } SYSTEM_INFO;
static inline void GetSystemInfo(SYSTEM_INFO *s) { s->dwNumberOfProcessors = (DWORD)sysconf(_SC_NPROCESSORS_ONLN); }
#endif
