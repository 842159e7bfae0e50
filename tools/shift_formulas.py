import random
P=(1<<64)-(1<<32)+1; M32=(1<<32)-1; EPS=M32; M64=(1<<64)-1
def sub_c(a,b):  # canonical sub as in field.cuh (single correction)
    assert a<P and b<P
    d=(a-b)&M64
    if a<b: d=(d-EPS)&M64
    assert d<P
    return d
def fA(x,s):
    assert 0<s<32
    x0,x1=x&M32,x>>32
    y0=(x0<<s)&M32; y1=((x1<<s)|(x0>>(32-s)))&M32; y2=x1>>(32-s)
    a=(y1<<32)|y0; c=(y2<<32)|((~y2)&M32)
    sm=a+c
    r=sm&M64
    if not (sm>>64):
        assert r>=EPS
        r-=EPS
    return r
def fB(x,s):
    assert 32<=s<64
    sp=s-32
    x0,x1=x&M32,x>>32
    if sp: y0=(x0<<sp)&M32; y1=((x1<<sp)|(x0>>(32-sp)))&M32; y2=x1>>(32-sp)
    else: y0,y1,y2=x0,x1,0
    S=y0+y1; S0=S&M32; c=S>>32
    hi=S0+c; assert hi<=M32
    a=hi<<32
    b=y1+y2+c
    return sub_c(a,b)
def fC(x,s):  # returns x*2^s for 64<=s<96 via  -(x*2^-k)
    assert 64<=s<96
    k=96-s
    x0=x&M32
    xh=x>>k
    v=(x0<<(32-k))&M32 if k<32 else x0
    a=xh+v; assert a<P
    b=v<<32; assert b<P
    return sub_c(b,a)
def ref(x,s): return x*pow(2,s,P)%P
random.seed(1)
edge=[0,1,2,P-1,P-2,EPS,EPS+1,1<<32,(1<<32)-1,(1<<63),(1<<63)-1,0xFFFFFFFF00000000,0xFFFFFFFE00000001,0xFFFFFFFEFFFFFFFF,0x00000000FFFFFFFF,0x0000000100000000]
vals=edge+[random.randrange(P) for _ in range(20000)]+[random.randrange(1<<33) for _ in range(2000)]+[P-1-random.randrange(1<<33) for _ in range(2000)]+[(random.randrange(1<<32)<<32) for _ in range(2000)]+[ (M32<<32)|0 ]
for s in range(1,96):
    f=fA if s<32 else fB if s<64 else fC
    for x in vals:
        if x>=P: continue
        r=f(x,s)
        assert r<P and r==ref(x,s),(s,hex(x),hex(r),hex(ref(x,s)))
print("all shift formulas exact and canonical for s=1..95 on",len(vals),"values")
