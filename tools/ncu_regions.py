import csv, subprocess, io, collections, sys
rep=sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
launches=[];cur=[];seen=set();f=None
for r in rows:
    if r and r[0]=="File Path":
        if r[1] in seen: launches.append(cur);cur=[];seen=set()
        seen.add(r[1]);f=r[1];continue
    if r and r[0]=="Line No": hdr=r;continue
    if r and r[0].isdigit() and len(r)>8: cur.append((f,r))
launches.append(cur)
ix={h:i for i,h in enumerate(hdr)}
src=open('ptsharp_b200/csrc/pt_device.cuh').read().split('\n')
def find(s): 
    for i,l in enumerate(src):
        if s in l: return i+1
marks={k:find(v) for k,v in dict(glue_start="} else if (nGlue > 0 && ", node_start="} else if (nNode > nLeaf) {", sched_start="const unsigned leafMask = __ballot_sync", march_start="if (nMarch > 0 && nMarch >= nGlue", boxi="PT_D void box_intersect(", prim_start="// Sphere.cs:40-60", tri_start="PT_D double triangle_intersect_exact", sdf_start="// Vector.LengthN", leafwork="PT_D void leaf_work(", rayaux="PT_D RayAux ray_aux", boxline="PT_D bool box_line_hit", meshstep="PT_D int mesh_step(", meshpop="PT_D bool mesh_pop(").items()}
for li,L in enumerate(launches):
    tot=sum(int(r[ix["Instructions Executed"]]) for _,r in L if r[ix["Instructions Executed"]].isdigit())
    cat=collections.Counter(); lanes=collections.Counter()
    for f,r in L:
        if not r[ix["Instructions Executed"]].isdigit(): continue
        n=int(r[ix["Instructions Executed"]]); t=int(r[ix["Thread Instructions Executed"]]); ln=int(r[0])
        if 'pt_device' not in f: c='other-file'
        elif ln<100: c='vecmath/netminmax'
        elif marks['boxi']<=ln<marks['prim_start']: c='box_intersect(FP64)'
        elif marks['prim_start']<=ln<marks['tri_start']: c='primitives'
        elif marks['tri_start']<=ln<marks['sdf_start']: c='triangle'
        elif marks['rayaux']<=ln<marks['rayaux']+8: c='ray_aux'
        elif marks['boxline']<=ln<marks['boxline']+12: c='box_line_hit'
        elif marks['meshpop']<=ln<marks['meshpop']+16: c='mesh_pop'
        elif marks['meshstep']<=ln<marks['leafwork']: c='mesh_step'
        elif marks['leafwork']<=ln<marks['leafwork']+35: c='leaf_work'
        elif marks['sched_start']<=ln<marks['march_start']: c='scheduler'
        elif marks['glue_start']<=ln<marks['node_start']: c='GLUE block'
        elif marks['node_start']<=ln<marks['node_start']+25: c='NODE/LEAF block'
        else: c='misc'
        cat[c]+=n; lanes[c]+=t
    print('launch',li,'warp inst',tot/1e6)
    for c,n in cat.most_common(): print(f"  {c:22s} {n/tot*100:5.1f}%  lanes {lanes[c]/max(n,1):5.1f}")
