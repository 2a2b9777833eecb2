"""Random darknet cfg / weights generator for the loader parity tests (tests/test_host.py).

Graphs are valid for the reference's parser (ffcnn.c:114-239): conv / depthwise / grouped conv with or without batch norm,
max and avg pools under both spellings, upsample, dropout, shortcut, routes written relative and absolute, yolo heads with
extra darknet keys; formatting varies (spaces around '=', shuffled keys, comment lines, blank lines, CRLF)."""
import numpy as np


def gen(rng):
    W=int(rng.choice([64,96,128,160])); H=int(rng.choice([64,96,128]))
    eq=lambda: rng.choice(["=", " = ", "= ", " ="])
    nl="\r\n" if rng.random()<0.15 else "\n"
    out=[]; convs=[]
    def sec(name, kv, shuffle=True):
        items=list(kv)
        if shuffle: rng.shuffle(items)
        lines=["[%s]"%name]
        for k,v in items:
            if rng.random()<0.1: lines.append("# note %d"%rng.integers(100))
            lines.append("%s%s%s"%(k,eq(),v))
        if rng.random()<0.5: lines.append("")
        out.append(nl.join(lines)+nl)
    sec("net",[("batch",1),("subdivisions",1),("width",W),("height",H),("channels",3),("momentum",0.9),("decay",0.0005),("learning_rate",0.001),("policy","steps"),("max_batches",1000)])
    shapes=[]   # output (w,h,c) per layer
    cur=(W,H,3)
    nl_layers=int(rng.integers(6,22)); yolo_done=0
    for li in range(nl_layers):
        w,h,c=cur
        r=rng.random()
        if r<0.5 or li==0:
            fs=int(rng.choice([1,3,3,5])); st=int(rng.choice([1,1,2])) if min(w,h)>=8 else 1
            dw = fs>1 and rng.random()<0.35
            grp = c if dw else (int(rng.choice([1,1,1,2,4])) if c%4==0 else 1)
            fn = c if dw else grp*int(rng.integers(1,9))*(1 if grp>1 else 4)
            bn=int(rng.random()<0.7); act=str(rng.choice(["leaky","relu","linear","mish"]))
            kv=[("filters",fn),("size",fs),("activation",act)]
            padv=int(rng.choice([1,1,1,0]))
            if rng.random()<0.9: kv.append(("pad",padv))
            else: padv=0
            if rng.random()<0.9 or st!=1: kv.append(("stride",st))
            if grp!=1 or rng.random()<0.2: kv.append(("groups",grp))
            if bn or rng.random()<0.3: kv.append(("batch_normalize",bn))
            pad=fs//2 if padv else 0
            ow=(w-fs+2*pad)//st+1; oh=(h-fs+2*pad)//st+1
            if ow<4 or oh<4: continue
            sec(str(rng.choice(["convolutional","conv"])),kv)
            convs.append((fn,fs,c//grp,bn)); cur=(ow,oh,fn)
        elif r<0.6 and min(w,h)>=8:
            fs=int(rng.choice([2,3,5,9])); st=int(rng.choice([1,2])) if (w%2==0 and h%2==0) else 1   # the reference's pools overrun their output on odd maps (ffcnn.c:381-394)
            sec(str(rng.choice(["maxpool","max","avgpool","avg"])),[("size",fs),("stride",st)]); cur=(w//st,h//st,c)
        elif r<0.65 and max(w,h)<=64:
            sec("upsample",[("stride",2)]); cur=(w*2,h*2,c)
        elif r<0.7:
            sec("dropout",[("probability",.15)])
        elif r<0.8:
            cand=[k for k in range(len(shapes)-1) if shapes[k]==cur]
            if not cand: continue
            k=int(rng.choice(cand)); sec("shortcut",[("from",k-len(shapes)),("activation",str(rng.choice(["linear","leaky"])))])
        elif r<0.9 and len(shapes)>=2:
            cand=[k for k in range(len(shapes)) if shapes[k][:2]==cur[:2]]
            n=int(rng.integers(1,min(4,len(cand))+1)); ks=[int(v) for v in rng.choice(cand,n,replace=False)]
            txt=[]; 
            for k in ks:
                txt.append(str(k) if (k>0 and rng.random()<0.4) else str(k-len(shapes)))
            sec("route",[("layers",str(rng.choice([", ",","])).join(txt))]); cur=(cur[0],cur[1],sum(shapes[k][2] for k in ks))
        else:
            if yolo_done>=2 or c<21: continue
            # head conv then yolo
            ncls=int(rng.integers(1,5)); fn=3*(5+ncls)
            sec("convolutional",[("filters",fn),("size",1),("stride",1),("pad",1),("activation","linear")]); convs.append((fn,1,c,0)); shapes.append((w,h,fn))
            masks=[int(v) for v in rng.choice(6,3,replace=False)]
            kv=[("mask",",".join(map(str,masks))),("anchors",", ".join("%d,%d"%(a,b) for a,b in rng.integers(4,90,(6,2)))),("classes",ncls),("num",6),("jitter",.3),("ignore_thresh",str(rng.choice([".3","0.5",".7"]))),("truth_thresh",1),("random",1)]
            if rng.random()<0.5: kv.append(("scale_x_y",str(rng.choice(["1.05","1.1","1.2"]))))
            sec("yolo",kv); yolo_done+=1
            shapes.append((0,0,0))                      # a yolo layer has no output tensor (ffcnn.c:190-211 leaves it 0x0x0)
            if li+2>=nl_layers: return "".join(out), convs, (W,H)
            k=int(rng.choice([j for j in range(len(shapes)-2) if shapes[j][2]>0]))   # darknet graphs continue with a route to an earlier layer
            sec("route",[("layers",str(k) if (k>0 and rng.random()<0.4) else str(k-len(shapes)))]); cur=shapes[k]
        shapes.append(cur)
    return "".join(out), convs, (W,H)

def weights(rng, convs, truncate=None):
    parts=[np.array([0,2,5],"<i4").tobytes(), np.array([1],"<u8").tobytes()]
    for fn,k,cpg,bn in convs:
        parts.append(rng.uniform(-.3,.3,fn).astype("<f4").tobytes())
        if bn:
            parts.append(rng.uniform(.6,1.4,fn).astype("<f4").tobytes()); parts.append(rng.uniform(-.2,.2,fn).astype("<f4").tobytes()); parts.append(rng.uniform(.3,1.2,fn).astype("<f4").tobytes())
        parts.append(rng.standard_normal(fn*cpg*k*k).astype("<f4").tobytes())
    b=b"".join(parts)
    return b if truncate is None else b[:int(len(b)*truncate)]

