import re, subprocess, sys
log=open(sys.argv[1] if len(sys.argv)>1 else '/tmp/nvcc_build.log').read()
pat=sys.argv[2] if len(sys.argv)>2 else '_f<'
blocks=re.split(r"ptxas info    : Compiling entry function '", log)[1:]
for b in blocks:
    name=b.split("'")[0]
    m=re.search(r"Used (\d+) registers", b)
    sp=re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
    dem=subprocess.run(['c++filt',name],capture_output=True,text=True).stdout.strip()
    short=re.sub(r"\(.*","",dem)
    if pat in short:
        print(short, 'regs',m.group(1) if m else None, 'stack/spill', sp.groups() if sp else None)
