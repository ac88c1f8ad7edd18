"""Development aid: lengths and instruction mix of the loops of the PEE kernels (from cuobjdump -sass of the built
object), to compare instruction counts of the step loops before spending GPU time.
usage: python scripts/sass_loops.py [object] ; python scripts/sass_loops.py dump <function regex> <first> <last>"""
import re,subprocess,sys
obj=sys.argv[1] if len(sys.argv)>1 else '/root/repo/codec_tcc_b200/lib/peeb_pee2.o'
txt=subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout
funcs=re.split(r'\n\s+Function : ',txt)
for f in funcs[1:]:
    name=f.split('\n',1)[0]
    if not re.search(r'ItLi256ELi3E',name): continue
    lines=[re.sub(r'\s*/\*.*$','',re.sub(r'^\s+/\*[0-9a-f]+\*/\s+','',l)) for l in f.split('\n') if re.match(r'^\s+/\*[0-9a-f]{4}\*/',l)]
    print(name[:40],'total',len(lines))
    for i,l in enumerate(lines):
        if 'BRA' in l:
            m=re.search(r'0x([0-9a-f]+)',l)
            if m:
                t=int(m.group(1),16)//16
                if t<i and i-t>60:
                    body=lines[t:i+1]
                    if sum('IDP' in b for b in body)==0: continue
                    # fast part: up to first '@!P0 BRA' after a VOTE following IDPs
                    print(f"  loop {t}->{i} len {i-t+1} IDP {sum('IDP' in b for b in body)} MOV {sum('MOV' in b for b in body)} UMOV {sum('UMOV' in b for b in body)} SEL {sum(b.startswith('SEL') for b in body)}")


def dump(pattern, first, last, obj=obj):
    """print instructions [first, last] of the first function whose name matches `pattern`"""
    for f in funcs[1:]:
        name = f.split('\n', 1)[0]
        if not re.search(pattern, name):
            continue
        ls = [re.sub(r'\s*/\*.*$', '', re.sub(r'^\s+/\*[0-9a-f]+\*/\s+', '', l)) for l in f.split('\n') if re.match(r'^\s+/\*[0-9a-f]{4}\*/', l)]
        for i in range(first, min(last, len(ls) - 1) + 1):
            print(i, hex(i * 16), ls[i])
        break
