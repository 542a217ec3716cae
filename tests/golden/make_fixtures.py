"""Regenerates the fixtures in this directory from the upstream reference tree (run where
/root/reference exists; the GPU box only ever sees the committed copies).

The reference holds exactly five data fixtures for the hot path: the two
Gmsh meshes, the reference's own numbering of bowl.msh (topology.dat) and the fields before /
after one Euler step of testGaussWave (out0.dat, out1.dat; examples/Main.cpp:172-195). They are
data (no source code) and are copied verbatim, gzip-compressed where large.
"""
import gzip
import os
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

PLAIN = {"examples/bowl.msh": "bowl.msh", "notebooks/basic.msh": "basic.msh"}
GZ = {"notebooks/topology.dat": "topology.dat.gz", "notebooks/out0.dat": "out0.dat.gz",
      "notebooks/out1.dat": "out1.dat.gz", "notebooks/geometry.dat": "geometry.dat.gz"}

for src, dst in PLAIN.items():
    shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, dst))
for src, dst in GZ.items():
    with open(os.path.join(REF, src), "rb") as f, open(os.path.join(HERE, dst), "wb") as raw:
        with gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as g:
            g.write(f.read())
print("fixtures written to", HERE)
