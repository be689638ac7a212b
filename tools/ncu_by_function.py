import re,collections,csv,sys
tag=sys.argv[1]  # e.g. Li10ELi10ELi5ELi32
srccsv=sys.argv[2]
# offset -> (func,file,line)
m={} ; inw=False; cur=None; lab=None
for line in open(sys.argv[3]):
    if line.startswith('\t.section\t.text.'):
        inw = tag in line
    if not inw: continue
    mm=re.search(r'//## File "([^"]+)", line (\d+)',line)
    if mm: cur=(mm.group(1).split('/')[-1],int(mm.group(2))); continue
    mm=re.match(r'(\$?[_A-Za-z0-9\$]+):',line)
    if mm and not mm.group(1).startswith('.L'):
        lab=mm.group(1); 
        lab=lab.split('$')[-1] if '$' in lab else 'main'
    mm=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S+)',line)
    if mm: m[int(mm.group(1),16)]=(lab,cur,mm.group(2))
allrows=list(csv.reader(open(srccsv)))
starts=[i for i,r in enumerate(allrows) if r and r[0]=='Kernel Name']+[len(allrows)]
want=tuple(re.findall(r'Li(\d+)E',tag))
sel=None
for a,b in zip(starts[:-1],starts[1:]):
    if tuple(re.findall(r'\(int\)(\d+)',allrows[a][1]))==want: sel=(a,b); break
if sel is None: sel=(starts[0],starts[1])
rows=allrows[sel[0]:sel[1]]
print(rows[0][1][:100])
hdr=rows[1]; 
ia=hdr.index('Address'); isamp=hdr.index('# Samples'); iex=hdr.index('Instructions Executed'); ith=hdr.index('Thread Instructions Executed')
base=None
byf=collections.Counter(); byfi=collections.Counter(); byl=collections.Counter(); byli=collections.Counter(); byth=collections.Counter()
tot=0;toti=0
for r in rows[2:]:
    a=int(r[ia],16)
    if base is None: base=a
    off=a-base
    f,cur,op=m.get(off,('?',None,'?'))
    s=int(r[isamp]); e=int(r[iex]); t=int(r[ith])
    tot+=s; toti+=e
    byf[f]+=s; byfi[f]+=e; byth[f]+=t
    key=(f if f!='main' else '', cur[0] if cur else '?', (cur[1]//10*10) if cur else 0)
    byl[key]+=s; byli[key]+=e
print('samples',tot,'inst',toti)
for f,v in byf.most_common(): print('%-40s samples %5.1f%%  inst %5.1f%%  lanes %.1f'%(f[-40:],100*v/tot,100*byfi[f]/toti, byth[f]/max(1,byfi[f])))
print('--- by line bucket (samples%, inst%)')
for k,v in byl.most_common(40): print('%5.1f%% %5.1f%%'%(100*v/tot,100*byli[k]/toti),k)
