import re,sys,collections
ALU={'IADD3','LOP3','SHF','PRMT','LEA','ISETP','SEL','VIADD','IABS','IMNMX','VIMNMX','FLO','POPC','BREV','MOV','CS2R','PLOP3','P2R','R2P','IADD','UIADD3','ULOP3','USHF'}
cur=None; h=collections.defaultdict(collections.Counter)
for l in open(sys.argv[1]):
    m=re.search(r'Function : (\S+)',l)
    if m: cur=m.group(1); continue
    m=re.search(r'^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)',l)
    if m and cur: h[cur][m.group(2)]+=1
for k,c in h.items():
    tot=sum(c.values()); alu=sum(v for o,v in c.items() if o.split('.')[0] in ALU); fma=sum(v for o,v in c.items() if o.startswith('IMAD') or o.startswith('FFMA') or o.startswith('FMUL') or o.startswith('HFMA2'))
    wide=sum(v for o,v in c.items() if o.startswith('IMAD.WIDE') or o.startswith('IMAD.HI'))
    print(k,'total',tot,'ALU',alu,'FMA',fma,'(wide/hi',wide,')')
    print('   ',', '.join(f'{o}:{v}' for o,v in c.most_common(18)))
