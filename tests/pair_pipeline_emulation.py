"""The dataflow of k_sha256_chains_pair (modelx_b200/csrc/sha256_kernels.cu) in plain Python: one SHA-256 chain run by
two lanes -- the E lane holds e,f,g,h, the A lane a,b,c,d and runs two rounds behind -- that exchange one value per
iteration (what shfl.xor does on the GPU) and use it one iteration later.  Each lane executes the same generic step
with per-lane constants.  Test infrastructure: tests/test_pair_pipeline.py checks it against hashlib, the GPU parity
tests check the kernel itself."""
import hashlib, struct, os
M=0xffffffff
K=[0x428a2f98,0x71374491,0xb5c0fbcf,0xe9b5dba5,0x3956c25b,0x59f111f1,0x923f82a4,0xab1c5ed5,0xd807aa98,0x12835b01,0x243185be,0x550c7dc3,0x72be5d74,0x80deb1fe,0x9bdc06a7,0xc19bf174,0xe49b69c1,0xefbe4786,0x0fc19dc6,0x240ca1cc,0x2de92c6f,0x4a7484aa,0x5cb0a9dc,0x76f988da,0x983e5152,0xa831c66d,0xb00327c8,0xbf597fc7,0xc6e00bf3,0xd5a79147,0x06ca6351,0x14292967,0x27b70a85,0x2e1b2138,0x4d2c6dfc,0x53380d13,0x650a7354,0x766a0abb,0x81c2c92e,0x92722c85,0xa2bfe8a1,0xa81a664b,0xc24b8b70,0xc76c51a3,0xd192e819,0xd6990624,0xf40e3585,0x106aa070,0x19a4c116,0x1e376c08,0x2748774c,0x34b0bcb5,0x391c0cb3,0x4ed8aa4a,0x5b9cca4f,0x682e6ff3,0x748f82ee,0x78a5636f,0x84c87814,0x8cc70208,0x90befffa,0xa4506ceb,0xbef9a3f7,0xc67178f2]
def rotr(x,n): return ((x>>n)|(x<<(32-n)))&M
def sched(block):
    w=list(struct.unpack(">16I",block))
    for t in range(16,64):
        s0=rotr(w[t-15],7)^rotr(w[t-15],18)^(w[t-15]>>3); s1=rotr(w[t-2],17)^rotr(w[t-2],19)^(w[t-2]>>10)
        w.append((w[t-16]+s0+w[t-7]+s1)&M)
    return [(w[t]+K[t])&M for t in range(64)]
def ch(x,y,z): return ((x&y)|(~x&z))&M
def compress_pair(H, block):
    wk=sched(block)
    # lanes: index 0 = E, 1 = A
    rot=[(6,11,25),(2,13,22)]; coef=[1,M]; mA=[0,M]
    hs=[H[4:8],H[0:4]]                       # E: H4..7 ; A: H0..3
    v=[[H[4],H[5],H[6],H[7]],[H[2],H[3],0xdeadbeef,0xdeadbeef]]
    ovr={0:H[1],1:H[0]}                      # A's `new` in the two peeled iterations
    # prologue: in1(0) for E = shfl(hs[3]) + hs[3]*coef + WK(0)
    recv=[hs[1][3],hs[0][3]]
    in1=[(recv[l]+(hs[l][3]*coef[l]+(wk[0] if l==0 else 0)))&M for l in range(2)]
    saved=None
    for i in range(66):
        send=[v[0][0],v[1][0]]; recv=[send[1],send[0]]
        new=[0,0]; in1n=[0,0]
        for l in range(2):
            v0,v1,v2,v3=v[l]
            X=rotr(v0,rot[l][0])^rotr(v0,rot[l][1])^rotr(v0,rot[l][2])
            p=(v1|(v2&mA[l]))&M; q=(v2&(v1|(~mA[l]&M)))&M
            C=ch(v0,p,q)
            wkx=(wk[i+1] if (l==0 and i+1<64) else 0)
            hWm=(v2*coef[l]+wkx)&M
            in1n[l]=(recv[l]+hWm)&M
            new[l]=(X+C+in1[l])&M
        if i in ovr: new[1]=ovr[i]
        for l in range(2): v[l]=[new[l]]+v[l][:3]
        in1=in1n
        if i==63: saved=list(v[0])
    v[0]=saved
    out=[0]*8
    for k in range(4):
        out[4+k]=(H[4+k]+v[0][k])&M; out[k]=(H[k]+v[1][k])&M
    return out
def sha256_pair(msg):
    H=[0x6a09e667,0xbb67ae85,0x3c6ef372,0xa54ff53a,0x510e527f,0x9b05688c,0x1f83d9ab,0x5be0cd19]
    ml=len(msg); msg=msg+b"\x80"+b"\0"*((55-ml)%64)+struct.pack(">Q",ml*8)
    for o in range(0,len(msg),64): H=compress_pair(H,msg[o:o+64])
    return struct.pack(">8I",*H)
if __name__ == "__main__":
    for n in [0,1,55,56,63,64,65,119,120,1000,4096]:
        m=os.urandom(n); assert sha256_pair(m)==hashlib.sha256(m).digest(),n
    print("pair pipeline emulation: ok")
