//go:build cgo && modelx_cuda

// SOURCE ONLY -- never compiled in the build environment of modelx-b200 (no Go toolchain there).
// Opt-in cgo binding of libmodelxdigest.so for kubegems/modelx.  Build the client with
//   CGO_ENABLED=1 go build -tags modelx_cuda ./cmd/modelx
// The default build (CGO_ENABLED=0, Makefile:63) keeps using digest_purego.go.
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

// digestFile replaces the body of Client.digest (pkg/client/push.go:149-161) and the hash in
// pullFile (pkg/client/pull.go:115-123): whole-file SHA-256, reference-identical result.
func digestFile(ctx context.Context, path string) (digest.Digest, error) {
	c, err := mxdOpen()
	if err != nil {
		return "", err
	}
	cpath := C.CString(path)
	defer C.free(unsafe.Pointer(cpath))
	var out [32]C.uint8_t
	var size C.uint64_t
	done := make(chan struct{})
	go func() { // ctx cancel: the reference closes the fd (push.go:156-159); here it aborts the stream
		select {
		case <-ctx.Done():
			C.mxd_cancel(c)
		case <-done:
		}
	}()
	rc := C.mxd_sha256_file(c, cpath, &out[0], &size)
	close(done)
	if rc != C.MXD_OK {
		if rc == C.MXD_ERR_CANCELED {
			C.mxd_reset_cancel(c)
			return "", ctx.Err()
		}
		return "", mxdError(rc)
	}
	var s [72]C.char
	C.mxd_digest_string(&out[0], &s[0])
	return digest.Digest(C.GoString(&s[0])), nil
}

// digestFiles hashes every blob of a push/pull in one lock-step GPU batch instead of
// PullPushConcurrency=3 goroutines (push.go:27,36-52; pull.go:41-50).
func digestFiles(paths []string) ([]digest.Digest, []int64, error) {
	c, err := mxdOpen()
	if err != nil {
		return nil, nil, err
	}
	n := len(paths)
	cpaths := make([]*C.char, n)
	for i, p := range paths {
		cpaths[i] = C.CString(p)
		defer C.free(unsafe.Pointer(cpaths[i]))
	}
	out := make([]C.uint8_t, 32*n)
	sizes := make([]C.uint64_t, n)
	if rc := C.mxd_sha256_files(c, (**C.char)(unsafe.Pointer(&cpaths[0])), C.uint64_t(n), &out[0], &sizes[0]); rc != C.MXD_OK {
		return nil, nil, mxdError(rc)
	}
	ds := make([]digest.Digest, n)
	sz := make([]int64, n)
	for i := range ds {
		var s [72]C.char
		C.mxd_digest_string(&out[32*i], &s[0])
		ds[i] = digest.Digest(C.GoString(&s[0]))
		sz[i] = int64(sizes[i])
	}
	return ds, sz, nil
}

// TreeDigest is the new chunked content address: chunk digests for the manifest annotation and
// the root, computed on every GPU the context drives (single process, chunk-range sharding).
func TreeDigest(path string) (root digest.Digest, chunks []digest.Digest, size int64, err error) {
	c, err := mxdOpen()
	if err != nil {
		return "", nil, 0, err
	}
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
	if rc := C.mxd_tree_digest_file(c, cpath, nil, &buf[0], C.uint64_t(capChunks), &n, &sz, &r[0]); rc != C.MXD_OK {
		return "", nil, 0, mxdError(rc)
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
