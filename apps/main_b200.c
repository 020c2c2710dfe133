/* main_b200.c — plain-C host program with the flow of the reference's main() (main.c:286-499) in which the
 * pthread fan-out over fill_canva (main.c:402-453) is ONE call into libtipe_rt.so (include/tipe_rt.h).
 *
 * Everything around that call stays host C, as in the reference: scene constants, OBJ/MTL + P3-PPM loading,
 * init_camera, the P3 writer.  The render knobs that main.c hard-codes (main.c:293-317) are command-line
 * options here, defaulting to the README scene (README.md:5-58), so configurations are data, not recompiles.
 *
 *   main_b200 [--scene readme|sky|none] [--obj F --mtl F] [--width W] [--ratio R] [--spp N] [--bounces N]
 *             [--ao 0|1] [--ao-intensity X] [--vfov D] [--origin x,y,z] [--target x,y,z] [--focus F]
 *             [--aperture ax,ay] [--seed S] [--gpus N] [--no-main-c-truncation] [--out file.ppm]
 *             [--tile-split]        several GPUs: interleaved 8x4 tiles per GPU instead of sample ranges (same image)
 *             [--sky-image F.ppm]   lat-long sky on the LAST sphere (the lookup main.c:64-71 keeps commented out; sky_file, main.c:322)
 *
 * Build: gcc -O2 apps/main_b200.c -o apps/main_b200 -Ltipe_raytracer_b200 -ltipe_rt -lm -Wl,-rpath,'$ORIGIN/../tipe_raytracer_b200'
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "trt_assets.h"

#define MAT(dr, dg, db, er, eg, eb, es, rs) {{{dr, dg, db}}, {{er, eg, eb}}, es, rs, 1.0, 1.0}

static double now_s(void) { struct timeval t; gettimeofday(&t, NULL); return (double)t.tv_sec + 1e-6 * (double)t.tv_usec; }

static int parse3(const char* s, double* out, int n)
{
    char* e = (char*)s;
    for (int k = 0; k < n; k++) { out[k] = strtod(e, &e); if (k + 1 < n) { if (*e != ',') return -1; e++; } }
    return 0;
}

int main(int argc, char** argv)
{
    /* ---- constants (main.c:293-317), README values ------------------------------------------------ */
    double ratio = 4.0 / 3.0; int largeur_image = 1000;
    double vfov = 70;
    double origin[3] = {0.34, 0.3, 0.5}, target[3] = {0.0, -0.5, -3}, up[3] = {0, 1, 0};
    double focus_distance = 3, ouverture[2] = {0.0, 0.0};
    int nbRayonParPixel = 100, nbRebondMax = 5;
    int useAO = 1; double AO_intensity = 2.5;
    const char *obj_file = NULL, *mtl_file = NULL, *scene = "readme", *out_file = "out_b200.ppm", *sky_file = NULL;
    unsigned long long seed = 1010; int n_gpus = 1, truncate_like_main_c = 1, tile_split = 0;

    for (int k = 1; k < argc; k++) {
        const char* a = argv[k]; const char* v = k + 1 < argc ? argv[k + 1] : NULL;
#define OPT(name) (strcmp(a, name) == 0 && v && (k++, 1))
        if (OPT("--scene")) scene = v; else if (OPT("--obj")) obj_file = v; else if (OPT("--mtl")) mtl_file = v;
        else if (OPT("--width")) largeur_image = atoi(v); else if (OPT("--ratio")) ratio = atof(v);
        else if (OPT("--spp")) nbRayonParPixel = atoi(v); else if (OPT("--bounces")) nbRebondMax = atoi(v);
        else if (OPT("--ao")) useAO = atoi(v); else if (OPT("--ao-intensity")) AO_intensity = atof(v);
        else if (OPT("--vfov")) vfov = atof(v); else if (OPT("--focus")) focus_distance = atof(v);
        else if (OPT("--origin")) { if (parse3(v, origin, 3)) goto usage; } else if (OPT("--target")) { if (parse3(v, target, 3)) goto usage; }
        else if (OPT("--aperture")) { if (parse3(v, ouverture, 2)) goto usage; }
        else if (OPT("--seed")) seed = strtoull(v, NULL, 10); else if (OPT("--gpus")) n_gpus = atoi(v); else if (OPT("--out")) out_file = v;
        else if (OPT("--sky-image")) sky_file = v;
        else if (strcmp(a, "--no-main-c-truncation") == 0) truncate_like_main_c = 0;
        else if (strcmp(a, "--tile-split") == 0) tile_split = 1;
        else { usage: fprintf(stderr, "usage: see the header of apps/main_b200.c (bad option `%s`)\n", a); return 2; }
    }
    int hauteur_image = (int)(largeur_image / ratio);

    /* ---- spheres (README.md:48-57 with alpha = materialIndex = 1; last sphere = sky, main.c:346) ------- */
    trt_sphere readme_spheres[] = {
        {{{-501, 0, 0}}, 500, MAT(0, 1, 0, 0, 0, 0, 0.0, 0.96)},
        {{{0, -501, 0}}, 500, MAT(1, 1, 1, 0, 0, 0, 0.0, 0.0)},
        {{{501, 0, 0}}, 500, MAT(1, 0, 0, 0, 0, 0, 0.0, 0.96)},
        {{{-0.5, 1.4, -1.2}}, 0.5, MAT(0, 0, 0, 1.0, 0.6, 0.2, 4.0, 0.0)},
        {{{0.5, 1.4, -2.2}}, 0.5, MAT(0, 0, 0, 0.7, 0.2, 1.0, 4.0, 0.0)},
        {{{0.6, -1.4, -1.0}}, 0.5, MAT(0, 0, 0, 0.55, 0.863, 1.0, 2.5, 0.0)},
        {{{-0.5, -1.4, -3.1}}, 0.5, MAT(0, 0, 0, 0.431, 1.0, 0.596, 2.5, 0.0)},
        {{{0, 0, -504}}, 500, MAT(1, 1, 1, 0, 0, 0, 0.0, 0.0)},
        {{{0, 501, 0}}, 500, MAT(1, 1, 1, 0, 0, 0, 0.0, 0.0)},
        {{{0.4, -0.5, -3.3}}, 0.5, MAT(0.784, 0.965, 1, 0, 0, 0, 0.0, 0.99)},
    };
    trt_sphere sky_sphere[] = {{{{0.0, 0.0, 0.0}}, 100000, MAT(0, 0, 0, 0.784, 0.965, 1, 1.0, 0.0)}};
    trt_sphere* sphere_list = NULL; int nbSpheres = 0;
    if (strcmp(scene, "readme") == 0) { sphere_list = readme_spheres; nbSpheres = 10; }
    else if (strcmp(scene, "sky") == 0) { sphere_list = sky_sphere; nbSpheres = 1; }
    else if (strcmp(scene, "none") != 0) { fprintf(stderr, "unknown --scene %s\n", scene); return 2; }

    double t_start = now_s();

    /* ---- mesh + textures (main.c:356-374) ----------------------------------------------------------- */
    char err[256] = "";
    trt_mesh mesh; memset(&mesh, 0, sizeof mesh);
    trt_material* mat_list = NULL; int tex_width = 0, tex_height = 0;
    if (obj_file) {
        if (!mtl_file) { fprintf(stderr, "--obj needs --mtl\n"); return 2; }
        if (trt_load_mesh(obj_file, mtl_file, &mesh, err) != 0) { fprintf(stderr, "%s\n", err); return 1; }
        trt_move_mesh(&mesh, 0, 0, 0);
        mat_list = trt_load_textures(mesh.texture_paths, mesh.nbMaterials, &tex_width, &tex_height, err);
        if (!mat_list) { fprintf(stderr, "%s\n", err); return 1; }   /* the reference would segfault at the first textured hit */
        printf("%s : %d triangles, %d textures of %dx%d loaded in %.3f s\n", obj_file, mesh.nbTriangles, mesh.nbMaterials,
               tex_width, tex_height, now_s() - t_start);
    }

    trt_camera cam = trt_init_camera(origin, target, up, vfov, ratio);
    size_t npix = (size_t)largeur_image * hauteur_image;
    trt_vec3* canva = (trt_vec3*)calloc(npix, sizeof(trt_vec3));
    trt_vec3* albedo_tab = (trt_vec3*)calloc(npix, sizeof(trt_vec3));
    trt_vec3* normal_tab = (trt_vec3*)calloc(npix, sizeof(trt_vec3));

    /* ---- the replacement for main.c:402-453 -------------------------------------------------------------- */
    trt_ctx* ctx = NULL;
    if (trt_init(NULL, n_gpus, &ctx) != TRT_OK) { fprintf(stderr, "trt_init: %s\n", trt_last_error(NULL)); return 1; }
    if (trt_scene_set(ctx, sphere_list, nbSpheres, mesh.triangles, mesh.nbTriangles, mat_list, tex_width, tex_height,
                      mesh.nbMaterials, mesh.quelMatPourTri, NULL, -1) != TRT_OK) {
        fprintf(stderr, "trt_scene_set: %s\n", trt_last_error(ctx)); return 1;
    }
    trt_material* sky_mat_list = NULL; int sky_width = 0, sky_height = 0;
    if (sky_file) {                                            /* create_mat_list(sky_file, ...), main.c:373-374 */
        sky_mat_list = trt_load_sky(sky_file, &sky_width, &sky_height, err);
        if (!sky_mat_list) { fprintf(stderr, "%s\n", err); return 1; }
        if (trt_sky_set(ctx, sky_mat_list, sky_width, sky_height) != TRT_OK) { fprintf(stderr, "trt_sky_set: %s\n", trt_last_error(ctx)); return 1; }
    }
    trt_params p; memset(&p, 0, sizeof p);
    p.largeur_image = largeur_image; p.hauteur_image = hauteur_image;
    p.nbRayonParPixel = nbRayonParPixel; p.nbRebondMax = nbRebondMax;
    p.useAO = useAO;
    /* struct ThreadData declares these four as int (main.c:42-43): main.c renders with the truncated values */
    p.AO_intensity = truncate_like_main_c ? (double)(int)AO_intensity : AO_intensity;
    p.focus_distance = truncate_like_main_c ? (double)(int)focus_distance : focus_distance;
    p.ouverture_x = truncate_like_main_c ? (double)(int)ouverture[0] : ouverture[0];
    p.ouverture_y = truncate_like_main_c ? (double)(int)ouverture[1] : ouverture[1];
    p.cam = cam; p.seed = seed;
    if (tile_split) p.flags |= TRT_FLAG_TILE_SPLIT;
    trt_stats st;
    double t_render = now_s();
    if (trt_render(ctx, &p, canva, albedo_tab, normal_tab, &st) != TRT_OK) { fprintf(stderr, "trt_render: %s\n", trt_last_error(ctx)); return 1; }
    t_render = now_s() - t_render;

    /* useDenoiser (main.c:455) stays an external post-pass on canva / albedo_tab / normal_tab */
    if (trt_write_ppm(out_file, canva, largeur_image, hauteur_image) != 0) { fprintf(stderr, "cannot write %s\n", out_file); return 1; }

    printf("%dx%d, %d spp, %d bounces, AO %s: %llu paths, %llu rays, trace %.2f ms, render call %.2f ms (%.1f Mrays/s), total %.3f s -> %s\n",
           largeur_image, hauteur_image, nbRayonParPixel, nbRebondMax, useAO ? "on" : "off",
           (unsigned long long)st.paths, (unsigned long long)st.rays, st.ms_trace, t_render * 1e3,
           (double)st.rays / t_render / 1e6, now_s() - t_start, out_file);

    trt_destroy(ctx);
    free(canva); free(albedo_tab); free(normal_tab); free(mat_list); free(sky_mat_list);
    trt_free_mesh(&mesh);
    return 0;
}
