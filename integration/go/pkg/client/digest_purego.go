//go:build !(cgo && modelx_cuda)

// SOURCE ONLY -- the default (pure Go) side of the digest seam; identical behaviour to the
// reference's Client.digest (pkg/client/push.go:149-161).
package client

import (
	"context"

	"github.com/opencontainers/go-digest"
)

func digestFile(ctx context.Context, path string) (digest.Digest, error) { return digestFileGo(ctx, path) }

// prefetchDigests is a no-op without the GPU engine: the 3 goroutines hash as they always did.
func prefetchDigests(ctx context.Context, paths []string) (context.Context, error) { return ctx, nil }
