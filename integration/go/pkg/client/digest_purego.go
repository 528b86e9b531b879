//go:build !(cgo && modelx_cuda)

// SOURCE ONLY -- the default (pure Go) side of the digest seam; identical behaviour to the
// reference's Client.digest (pkg/client/push.go:149-161).
package client

import (
	"context"
	"os"

	"github.com/opencontainers/go-digest"
)

func digestFile(ctx context.Context, path string) (digest.Digest, error) {
	f, err := os.Open(path)
	if err != nil {
		return "", err
	}
	defer f.Close()
	go func() {
		<-ctx.Done()
		f.Close()
	}()
	return digest.FromReader(f)
}
