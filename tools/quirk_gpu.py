import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
import oracle as _o
from sregex_b200 import capi, cuda as cu
o=capi.load("oracle")
subs=[b'.b.a', b'.a.a', b'-1-2', b' ab cd', b'b.a', b'..ab.ab', b'x.b.1.b a']
pitch=32
for rx in (rb'\w+x?', rb'a+b?', rb'\d+\.?', rb'[a-z]+(\d)?', rb'(\w+) ?', rb'(\w+)+(\w+)?'):
    prog=cu.CudaProgram(rx); po=o.compile(rx,0)
    for tier in (0,1,2,3):
        prog.set_pike_tier(tier)
        bad=[]
        for s in subs:
            host=np.zeros((1,pitch),dtype=np.uint8); host[0,:len(s)]=np.frombuffer(s,dtype=np.uint8)
            rc,ov=prog.pike_lines(torch.from_numpy(host).cuda(),1,pitch,len(s))
            want=o.pike(po,s)
            got=(int(rc[0]), ov[0].cpu().tolist() if int(rc[0])>=0 else None)
            if got!=want: bad.append((s,got,want))
        print(rx, 'tier', tier, 'last', prog.last_pike_tier(), 'BAD' if bad else 'ok', bad[:2])
    # classic API
    bad=[]
    for s in subs:
        got=cu.pike(po.__class__ and prog.program, s, [s]) if hasattr(cu,'pike') else None
    prog.program.close()
