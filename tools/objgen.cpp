// tools/objgen.cpp — host-only helper (links the host front-end sources, no CUDA, not the product library): writes the
// bench's procedural meshes as Wavefront OBJ so that the REFERENCE renderer can load them (bench.py --impl reference must
// not load libhexray_b200.so).   usage: hxr_objgen terrain <grid side> <seed> out.obj | hxr_objgen soup <triangles> <seed> out.obj
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../hexray_b200/csrc/host/scene.h"

int main(int argc, char** argv)
{
    if (argc != 5) { fprintf(stderr, "usage: %s terrain|soup <n> <seed> out.obj\n", argv[0]); return 2; }
    hxr::host::Mesh m;
    const long long n = atoll(argv[2]);
    const unsigned long long seed = strtoull(argv[3], nullptr, 0);
    if (!strcmp(argv[1], "terrain")) m.generateTerrain((int)n, seed);
    else if (!strcmp(argv[1], "soup")) m.generateSoup(n, seed);
    else { fprintf(stderr, "unknown mesh kind %s\n", argv[1]); return 2; }
    if (!m.saveOBJ(argv[4])) { fprintf(stderr, "cannot write %s\n", argv[4]); return 1; }
    return 0;
}
