import re,collections,sys
kern=sys.argv[1]; src=sys.argv[2]
bounds=eval(sys.argv[3])
cur=None; fn=None
cnt=collections.Counter()
for line in open(sys.argv[4] if len(sys.argv)>4 else 'swin.dis'):
    m=re.match(r'\.text\.(\S+):',line)
    if m: fn=m.group(1); cur=None; continue
    if fn is None or kern not in fn: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)',line)
    if m:
        if m.group(1).endswith(src): cur=int(m.group(2))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,6}\*/',line):
        cnt[cur]+=1
tot=sum(cnt.values())
print('total',tot)
for lo,hi,name in bounds:
    n=sum(v for k,v in cnt.items() if k and lo<=k<=hi)
    print(f'{name:16s} {n:5d}')
lo0=bounds[0][0]; hi0=bounds[-1][1]
print('other', sum(v for k,v in cnt.items() if not k or k<lo0 or k>hi0))
for k,v in sorted(cnt.items(), key=lambda kv:-kv[1])[:int(sys.argv[5]) if len(sys.argv)>5 else 30]:
    print(k,v)
