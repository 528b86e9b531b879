"""ThreadSanitizer target: a Python-driven mixed load (tree digests, whole-file jobs, hashers, per-operation cancels, 5 threads) on\nthe -fsanitize=thread build of the CPU test double; run by tools/tsan_digest_service.sh with libtsan preloaded."""
import sys, os, random, hashlib, threading, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelx_b200
lib=sys.argv[1]
tmp=tempfile.mkdtemp(); rng=random.Random(7)
blobs=[rng.randbytes(rng.randrange(1,2_000_000)) for _ in range(5)]; paths=[]
for i,b in enumerate(blobs):
    p=os.path.join(tmp,f"m{i}"); open(p,"wb").write(b); paths.append(p)
want=[hashlib.sha256(b).digest() for b in blobs]; errs=[]
with modelx_b200.Engine(devices=[0], ring_bytes=8<<20, lib_path=lib) as eng:
    wt=[eng.tree_digest(b,1<<20,16<<10,8)[1] for b in blobs]
    def worker(seed):
        r=random.Random(seed)
        try:
            for _ in range(8):
                i=r.randrange(len(blobs)); k=r.randrange(4)
                if k==0: assert eng.tree_digest_file(paths[i],1<<20,16<<10,8)[1]==wt[i]
                elif k==1: assert eng.sha256_files([paths[i]])[0]==[want[i]]
                elif k==2:
                    h=eng.hasher(); h.write(blobs[i]); assert h.sum()==want[i]; h.close()
                else:
                    with eng.op() as op:
                        t=threading.Timer(r.random()*0.01, op.cancel); t.start()
                        try: assert op.sha256_file(paths[i])[0]==want[i]
                        except modelx_b200.MxdError as e: assert e.status==-6
                        t.join()
        except Exception as e: errs.append(repr(e))
    ts=[threading.Thread(target=worker,args=(s,)) for s in range(5)]
    [t.start() for t in ts]; [t.join() for t in ts]
print("errors",errs)
