import sys, numpy as np
sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/tests")
from oracle import datasets, reference_pipeline as rp
from databricks_kubernetes_mlops_poc_b200 import flatten, encode
from databricks_kubernetes_mlops_poc_b200.engine import ForestEngine
from blob_walk import walk_blob
cur=datasets.load_curated()
for params in (dict(n_estimators=1,max_depth=2,random_state=0), dict(n_estimators=1,max_depth=6,random_state=0), dict(n_estimators=32,max_depth=6,random_state=0)):
    pipe=rp.fit_reference_pipeline(cur.iloc[:5000], params)
    flat=flatten.flatten_pipeline(pipe); enc=encode.RowEncoder(flat)
    sub=cur.iloc[:2000]; rows=enc.encode_frame(sub)
    eng=ForestEngine(flat,0)
    p,l=eng.predict_rows(rows,np.float64)
    pe,le=walk_blob(flat.blob,rows)
    bad=np.nonzero(np.abs(p-pe)>1e-12)[0]
    print(params, "bad", len(bad), "of", len(p), "maxerr", np.abs(p-pe).max())
    if len(bad):
        i=bad[0]; print("row",i, rows[i], rows[i].view(np.float32)[9:23], "gpu",p[i],"emu",pe[i])
        h=flatten.parse_header(flat.blob); g=h["groups"][0]
        buf=np.frombuffer(flat.blob,dtype=np.uint8); base=h["chunks_off"]
        N=buf[base:base+g["n_slots"]*256].view(np.uint32).reshape(g["n_slots"],32,2)
        node=0
        w=rows[i].copy(); w[23]=0xffffffff
        for d in range(g["depth"]):
            t,m=N[node,0]; x=w[m&31]
            print(" d",d,"node",node,"feat",m&31,"cat",bool(m&32),"T",hex(t),np.uint32(t).view(np.float32),"x",hex(x),np.uint32(x).view(np.float32),"first",(m>>8))
            geu=not (np.uint32(x).view(np.float32) < np.uint32(t).view(np.float32))
            second=(x==t) or (geu and not (m&32))
            node=(m>>8)+int(second)
    eng.close()
