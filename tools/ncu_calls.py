"""usage: ncu_calls.py report.ncu-rep kernel_regex mangled_name_regex [matching .so]
Lists the executed CALL instructions of a kernel (slow-path subroutines of div/sqrt/rcp, ...) with their counts, active lanes and
source lines: the SASS page of an ncu report joined with nvdisasm line info of the .so the report was captured from."""
import sys, csv, subprocess, io
sys.path.insert(0,'/root/repo/tools')
import importlib.util
spec=importlib.util.spec_from_file_location("h","/root/repo/tools/ncu_hotspots.py"); h=importlib.util.module_from_spec(spec); spec.loader.exec_module(h)
from pathlib import Path
rep, kre, disre = sys.argv[1], sys.argv[2], sys.argv[3]
txt=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+kre],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(txt)))
hi=[i for i,r in enumerate(rows) if r and r[0]=="Address"][0]
hdr=rows[hi]; ci=hdr.index("Instructions Executed"); ti=hdr.index("Thread Instructions Executed"); si=hdr.index("# Samples")
dis=h.disasm(Path("sys.argv[4] if len(sys.argv) > 4 else "/root/repo/rs_pbrt_b200/librs_pbrt_b200.so""), disre)
data=rows[hi+1:hi+1+len(dis)]
tot=sum(int(d[ci]) for d in data)
calls=[]; sub=0
for i,(addr,t,frames) in enumerate(dis):
    e=int(data[i][ci])
    if t.startswith("CALL") and e>0: calls.append((e, round(int(data[i][ti])/max(e,1),1), t[22:70], frames[0] if frames else None, frames[-1] if frames else None))
calls.sort(reverse=True)
print("total warp-instr", tot, "; call sites executed:", len(calls), "calls", sum(c[0] for c in calls))
for c in calls[:14]: print(c)
