import numpy as np, time, os, tempfile, sys
sys.path.insert(0, os.getcwd())
from lammps_analysis_b200.synthetic import nacl_trajectory
from lammps_analysis_b200.file_io import write_lammps_dump, LAMMPSTrajectoryFile
data, box = nacl_trajectory(1000, 300, 32.0, seed=1)
tmp=tempfile.mkdtemp()
p=os.path.join(tmp,"t.lammpstraj")
write_lammps_dump(p, data, box, step_stride=10)
sz=os.path.getsize(p)/1e6
f=LAMMPSTrajectoryFile(p, native=True); f.metadata
for thr in ("1","4","16"):
    os.environ["MDK_INGEST_THREADS"]=thr
    t0=time.perf_counter(); off=0; done=0
    while done<300:
        k=min(64,300-done); blk,off=f._read_native(off,k,f._n_atoms,f._columns); done+=k
    t1=time.perf_counter()
    print("threads",thr,"MB/s",round(sz/(t1-t0)), flush=True)
print("cpus", os.cpu_count())
