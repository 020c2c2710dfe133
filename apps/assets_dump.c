/* assets_dump.c — test helper: loads an OBJ/MTL + PPM set with trt_assets.h and dumps the raw arrays
 * (triangle[], quelMatPourTri[], material[]) to a binary file, so tests can compare them byte for byte with
 * what the reference's list_of_mesh / create_mat_list_mtl produce.  No GPU involved. */
#include "trt_assets.h"

int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: assets_dump scene.obj scene.mtl out.bin\n"); return 2; }
    char err[256] = "";
    trt_mesh m;
    if (trt_load_mesh(argv[1], argv[2], &m, err) != 0) { fprintf(stderr, "%s\n", err); return 1; }
    int tw = 0, th = 0;
    trt_material* mats = trt_load_textures(m.texture_paths, m.nbMaterials, &tw, &th, err);
    if (!mats) { fprintf(stderr, "%s\n", err); return 1; }
    FILE* f = fopen(argv[3], "wb");
    if (!f) return 1;
    int hdr[4] = {m.nbTriangles, m.nbMaterials, tw, th};
    fwrite(hdr, sizeof hdr, 1, f);
    fwrite(m.triangles, sizeof(trt_triangle), (size_t)m.nbTriangles, f);
    fwrite(m.quelMatPourTri, sizeof(int), (size_t)m.nbTriangles, f);
    fwrite(mats, sizeof(trt_material), (size_t)m.nbMaterials * tw * th, f);
    fclose(f);
    for (int k = 0; k < m.nbMaterials; k++) printf("%s\n", m.texture_paths[k]);
    free(mats); trt_free_mesh(&m);
    return 0;
}
