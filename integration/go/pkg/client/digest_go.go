// SOURCE ONLY -- the reference's own digest (pkg/client/push.go:149-161), shared by both builds of the seam.
package client

import (
	"context"
	"os"

	"github.com/opencontainers/go-digest"
)

func digestFileGo(ctx context.Context, path string) (digest.Digest, error) {
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
