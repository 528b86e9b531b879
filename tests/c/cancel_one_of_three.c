/* C-level proof of the cancellation contract the cgo shim relies on (integration/go/pkg/client/digest_cuda.go):
 * three threads -- the three goroutines of PullPushConcurrency (push.go:27) -- each hash one file through its own
 * operation handle; the main thread cancels ONE of them (push.go:150-159: that call's ctx).  The canceled call must
 * return MXD_ERR_CANCELED (or finish with the right digest if it won the race), never a digest with status OK that
 * is wrong, and never "" with nil; the other two must finish with correct digests; nothing sticks to the context.
 *
 * usage: cancel_one_of_three <libmodelxdigest.so> <file0> <hex0> <file1> <hex1> <file2> <hex2>
 * Plain C, dlopen only: exactly what a non-Python host sees of the ABI. */
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include "../../include/modelx_digest.h"

static int (*p_open)(mxd_ctx**, const int*, int, uint64_t);
static void (*p_close)(mxd_ctx*);
static int (*p_op_begin)(mxd_ctx*, mxd_ctx**);
static void (*p_op_end)(mxd_ctx*);
static void (*p_cancel)(mxd_ctx*);
static int (*p_sha256_file)(mxd_ctx*, const char*, uint8_t*, uint64_t*);
static void (*p_digest_string)(const uint8_t*, char*);

struct job { mxd_ctx* op; const char* path; int rc; char got[72]; };

static void* work(void* arg) {
    struct job* j = (struct job*)arg;
    uint8_t d[32]; uint64_t size = 0;
    j->got[0] = 0;
    j->rc = p_sha256_file(j->op, j->path, d, &size);
    if (j->rc == MXD_OK) p_digest_string(d, j->got);
    return NULL;
}

int main(int argc, char** argv) {
    if (argc != 8) { fprintf(stderr, "usage: %s lib f0 hex0 f1 hex1 f2 hex2\n", argv[0]); return 2; }
    void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
#define SYM(p, name) do { *(void**)(&p) = dlsym(lib, name); if (!p) { fprintf(stderr, "missing %s\n", name); return 2; } } while (0)
    SYM(p_open, "mxd_open"); SYM(p_close, "mxd_close"); SYM(p_op_begin, "mxd_op_begin"); SYM(p_op_end, "mxd_op_end");
    SYM(p_cancel, "mxd_cancel"); SYM(p_sha256_file, "mxd_sha256_file"); SYM(p_digest_string, "mxd_digest_string");
    mxd_ctx* ctx = NULL;
    int dev0 = 0;
    int rc = p_open(&ctx, &dev0, 1, 4u << 20);
    if (rc != MXD_OK) { fprintf(stderr, "mxd_open: %d\n", rc); return 3; }
    int bad = 0;
    for (int round = 0; round < 3; ++round) {          /* cancel a different victim each round */
        struct job jobs[3]; pthread_t th[3];
        for (int i = 0; i < 3; ++i) { jobs[i].path = argv[2 + 2 * i]; if (p_op_begin(ctx, &jobs[i].op) != MXD_OK) return 3; }
        for (int i = 0; i < 3; ++i) pthread_create(&th[i], NULL, work, &jobs[i]);
        usleep(1500 + 700 * round);
        p_cancel(jobs[round].op);
        for (int i = 0; i < 3; ++i) pthread_join(th[i], NULL);
        for (int i = 0; i < 3; ++i) {
            char want[80]; snprintf(want, sizeof want, "sha256:%s", argv[3 + 2 * i]);
            if (i == round) {
                if (jobs[i].rc == MXD_ERR_CANCELED) { if (jobs[i].got[0]) { fprintf(stderr, "round %d: canceled call produced a digest\n", round); bad = 1; } }
                else if (jobs[i].rc != MXD_OK || strcmp(jobs[i].got, want)) { fprintf(stderr, "round %d: victim rc=%d digest=%s\n", round, jobs[i].rc, jobs[i].got); bad = 1; }
            } else if (jobs[i].rc != MXD_OK || strcmp(jobs[i].got, want)) {
                fprintf(stderr, "round %d: bystander %d rc=%d digest=%s want %s\n", round, i, jobs[i].rc, jobs[i].got, want); bad = 1;
            }
            printf("round %d call %d rc=%d %s\n", round, i, jobs[i].rc, i == round ? "(canceled)" : "");
            p_op_end(jobs[i].op);
        }
    }
    /* nothing sticks: a fresh call on the root handle works */
    uint8_t d[32]; uint64_t size = 0; char got[72], want[80];
    rc = p_sha256_file(ctx, argv[2], d, &size);
    p_digest_string(d, got); snprintf(want, sizeof want, "sha256:%s", argv[3]);
    if (rc != MXD_OK || strcmp(got, want)) { fprintf(stderr, "after the cancels: rc=%d %s\n", rc, got); bad = 1; }
    p_close(ctx);
    puts(bad ? "FAIL" : "PASS");
    return bad;
}
