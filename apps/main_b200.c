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
 *             [--no-overrides]      drop the three per-material overrides texture.h:71-87 hard-codes (ordinals 1, 4, 3)
 *             [--override ORD:field=value[,field=value...]]   set / replace the override of material ordinal ORD; fields:
 *                                   alpha, index, reflection, strength (emissionStrength), emission=r/g/b.  Repeatable.
 *             [--sphere "cx cy cz r  dr dg db  er eg eb  es rs alpha index"]   append a sphere (main.c:330-347); --scene none starts empty
 *             [--config FILE]       a scene file: one `option = value` per line with the option names above (no dashes), `#` comments;
 *                                   the constants of main.c:293-347 become data.  Later command-line options win.
 *             [--dry-run]           print the resolved configuration and exit before touching the GPU
 *             [--pageable-output]   canva / albedo_tab / normal_tab from calloc like main.c:385-400 (default: page-locked, trt_host_alloc)
 *             [--nranks N --rank R --id-file PATH [--device D]]   ONE PROCESS PER GPU: start N copies of this program with ranks
 *                                   0..N-1 and the same options; rank 0 writes the NCCL id to PATH, the others read it; every rank traces
 *                                   its share of the samples on GPU D (default R), the sums meet on rank 0 with ncclReduce, rank 0 writes the image
 *
 * Build: gcc -O2 apps/main_b200.c -o apps/main_b200 -Ltipe_raytracer_b200 -ltipe_rt -lm -Wl,-rpath,'$ORIGIN/../tipe_raytracer_b200'
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <time.h>

#include "trt_assets.h"

#define MAT(dr, dg, db, er, eg, eb, es, rs) {{{dr, dg, db}}, {{er, eg, eb}}, es, rs, 1.0, 1.0}

static double now_s(void) { struct timeval t; gettimeofday(&t, NULL); return (double)t.tv_sec + 1e-6 * (double)t.tv_usec; }

static int parse3(const char* s, double* out, int n)
{
    char* e = (char*)s;
    for (int k = 0; k < n; k++) { out[k] = strtod(e, &e); if (k + 1 < n) { if (*e != ',') return -1; e++; } }
    return 0;
}

/* ---- configuration: main.c:293-347 as data ------------------------------------------------------------------ */
#define MAX_SPHERES 1024
#define MAX_OVERRIDES 64
typedef struct config {
    double ratio; int largeur_image; double vfov;
    double origin[3], target[3], up[3];
    double focus_distance, ouverture[2];
    int nbRayonParPixel, nbRebondMax, useAO; double AO_intensity;
    const char *obj_file, *mtl_file, *out_file, *sky_file;
    unsigned long long seed; int n_gpus, truncate_like_main_c, tile_split, dry_run;
    int n_ranks, rank, device; const char* id_file;
    int pageable_output;
    trt_sphere spheres[MAX_SPHERES]; int nbSpheres;
    trt_mat_override overrides[MAX_OVERRIDES]; int n_overrides;      /* starts as texture.h:71-87 */
} config;

static const trt_sphere README_SPHERES[10] = {                        /* README.md:48-57 with alpha = materialIndex = 1 */
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
static const trt_sphere SKY_SPHERE = {{{0.0, 0.0, 0.0}}, 100000, MAT(0, 0, 0, 0.784, 0.965, 1, 1.0, 0.0)};   /* main.c:346 */

static void config_defaults(config* c)
{
    memset(c, 0, sizeof *c);
    c->ratio = 4.0 / 3.0; c->largeur_image = 1000; c->vfov = 70;
    c->origin[0] = 0.34; c->origin[1] = 0.3; c->origin[2] = 0.5;
    c->target[0] = 0.0; c->target[1] = -0.5; c->target[2] = -3; c->up[1] = 1;
    c->focus_distance = 3;
    c->nbRayonParPixel = 100; c->nbRebondMax = 5; c->useAO = 1; c->AO_intensity = 2.5;
    c->out_file = "out_b200.ppm"; c->seed = 1010; c->n_gpus = 1; c->truncate_like_main_c = 1;
    c->n_ranks = 1; c->rank = 0; c->device = -1;
    memcpy(c->spheres, README_SPHERES, sizeof README_SPHERES); c->nbSpheres = 10;
    /* same table as trt_default_overrides() (texture.h:71-87), written out so that --dry-run needs no library call */
    c->n_overrides = 3;
    c->overrides[0].material_ordinal = 1; c->overrides[0].mask = TRT_OV_EMISSION_COLOR | TRT_OV_EMISSION_STRENGTH | TRT_OV_ALPHA;
    c->overrides[0].emissionColor[0] = c->overrides[0].emissionColor[1] = c->overrides[0].emissionColor[2] = 1.0;
    c->overrides[0].emissionStrength = 1.85; c->overrides[0].alpha = 1.0;
    c->overrides[1].material_ordinal = 4; c->overrides[1].mask = TRT_OV_ALPHA | TRT_OV_MATERIAL_INDEX | TRT_OV_REFLECTION;
    c->overrides[1].alpha = 0.6; c->overrides[1].materialIndex = 1.33; c->overrides[1].reflectionStrength = 0.93;
    c->overrides[2].material_ordinal = 3; c->overrides[2].mask = TRT_OV_ALPHA | TRT_OV_MATERIAL_INDEX | TRT_OV_REFLECTION;
    c->overrides[2].alpha = 0.1; c->overrides[2].materialIndex = 1.50; c->overrides[2].reflectionStrength = 0.3;
}

/* "ORD:field=value,field=value": merged into the entry of that ordinal (created if absent) */
static int parse_override(config* c, const char* v)
{
    char buf[256]; strncpy(buf, v, sizeof buf - 1); buf[sizeof buf - 1] = 0;
    char* colon = strchr(buf, ':');
    if (!colon) return -1;
    *colon = 0;
    const int ord = atoi(buf);
    trt_mat_override* o = NULL;
    for (int k = 0; k < c->n_overrides; k++) if (c->overrides[k].material_ordinal == ord) o = &c->overrides[k];
    if (!o) {
        if (c->n_overrides >= MAX_OVERRIDES) return -1;
        o = &c->overrides[c->n_overrides++]; memset(o, 0, sizeof *o); o->material_ordinal = ord;
    }
    for (char* tok = strtok(colon + 1, ","); tok; tok = strtok(NULL, ",")) {
        char* eq = strchr(tok, '=');
        if (!eq) return -1;
        *eq = 0;
        const char* val = eq + 1;
        if (strcmp(tok, "alpha") == 0) { o->alpha = atof(val); o->mask |= TRT_OV_ALPHA; }
        else if (strcmp(tok, "index") == 0) { o->materialIndex = atof(val); o->mask |= TRT_OV_MATERIAL_INDEX; }
        else if (strcmp(tok, "reflection") == 0) { o->reflectionStrength = atof(val); o->mask |= TRT_OV_REFLECTION; }
        else if (strcmp(tok, "strength") == 0) { o->emissionStrength = atof(val); o->mask |= TRT_OV_EMISSION_STRENGTH; }
        else if (strcmp(tok, "emission") == 0) {
            if (sscanf(val, "%lf/%lf/%lf", &o->emissionColor[0], &o->emissionColor[1], &o->emissionColor[2]) != 3) return -1;
            o->mask |= TRT_OV_EMISSION_COLOR;
        } else return -1;
    }
    return 0;
}

static int config_load_file(config* c, const char* path, int depth);

/* one option (name without dashes) with its value; returns 0, or -1 on a bad option / value */
static int config_set(config* c, const char* name, const char* v, int depth)
{
#define IS(n) (strcmp(name, n) == 0)
    if (IS("no-main-c-truncation")) { c->truncate_like_main_c = 0; return 0; }
    if (IS("tile-split")) { c->tile_split = 1; return 0; }
    if (IS("no-overrides")) { c->n_overrides = 0; return 0; }
    if (IS("dry-run")) { c->dry_run = 1; return 0; }
    if (IS("pageable-output")) { c->pageable_output = 1; return 0; }
    if (!v) return -1;
    if (IS("scene")) {
        if (strcmp(v, "readme") == 0) { memcpy(c->spheres, README_SPHERES, sizeof README_SPHERES); c->nbSpheres = 10; }
        else if (strcmp(v, "sky") == 0) { c->spheres[0] = SKY_SPHERE; c->nbSpheres = 1; }
        else if (strcmp(v, "none") == 0) c->nbSpheres = 0;
        else return -1;
    }
    else if (IS("sphere")) {
        if (c->nbSpheres >= MAX_SPHERES) return -1;
        trt_sphere* s = &c->spheres[c->nbSpheres];
        memset(s, 0, sizeof *s);
        if (sscanf(v, "%lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf", &s->center.e[0], &s->center.e[1], &s->center.e[2], &s->radius,
                   &s->mat.diffuseColor.e[0], &s->mat.diffuseColor.e[1], &s->mat.diffuseColor.e[2],
                   &s->mat.emissionColor.e[0], &s->mat.emissionColor.e[1], &s->mat.emissionColor.e[2],
                   &s->mat.emissionStrength, &s->mat.reflectionStrength, &s->mat.alpha, &s->mat.materialIndex) != 14) return -1;
        c->nbSpheres++;
    }
    else if (IS("obj")) c->obj_file = strdup(v); else if (IS("mtl")) c->mtl_file = strdup(v);
    else if (IS("out")) c->out_file = strdup(v); else if (IS("sky-image")) c->sky_file = strdup(v);
    else if (IS("width")) c->largeur_image = atoi(v); else if (IS("ratio")) c->ratio = atof(v);
    else if (IS("spp")) c->nbRayonParPixel = atoi(v); else if (IS("bounces")) c->nbRebondMax = atoi(v);
    else if (IS("ao")) c->useAO = atoi(v); else if (IS("ao-intensity")) c->AO_intensity = atof(v);
    else if (IS("vfov")) c->vfov = atof(v); else if (IS("focus")) c->focus_distance = atof(v);
    else if (IS("origin")) return parse3(v, c->origin, 3); else if (IS("target")) return parse3(v, c->target, 3);
    else if (IS("up")) return parse3(v, c->up, 3);
    else if (IS("aperture")) return parse3(v, c->ouverture, 2);
    else if (IS("seed")) c->seed = strtoull(v, NULL, 10); else if (IS("gpus")) c->n_gpus = atoi(v);
    else if (IS("nranks")) c->n_ranks = atoi(v); else if (IS("rank")) c->rank = atoi(v); else if (IS("device")) c->device = atoi(v);
    else if (IS("id-file")) c->id_file = strdup(v);
    else if (IS("override")) return parse_override(c, v);
    else if (IS("config")) return config_load_file(c, v, depth + 1);
    else return -1;
    return 0;
#undef IS
}

static int is_flag(const char* name)
{
    return strcmp(name, "no-main-c-truncation") == 0 || strcmp(name, "tile-split") == 0 || strcmp(name, "no-overrides") == 0 || strcmp(name, "dry-run") == 0 || strcmp(name, "pageable-output") == 0;
}

static int config_load_file(config* c, const char* path, int depth)
{
    if (depth > 4) { fprintf(stderr, "%s: config files nested too deep\n", path); return -1; }
    FILE* f = fopen(path, "r");
    if (!f) { fprintf(stderr, "cannot open config %s\n", path); return -1; }
    char line[1024]; int no = 0, rc = 0;
    while (rc == 0 && fgets(line, sizeof line, f)) {
        no++;
        char* hash = strchr(line, '#'); if (hash) *hash = 0;
        char* b = line; while (*b == ' ' || *b == '\t') b++;
        char* e = b + strlen(b); while (e > b && (e[-1] == '\n' || e[-1] == '\r' || e[-1] == ' ' || e[-1] == '\t')) *--e = 0;
        if (!*b) continue;
        char* eq = strchr(b, '=');
        char* val = NULL;
        if (eq) { *eq = 0; val = eq + 1; while (*val == ' ' || *val == '\t') val++; }
        e = b + strlen(b); while (e > b && (e[-1] == ' ' || e[-1] == '\t')) *--e = 0;
        if (config_set(c, b, is_flag(b) ? NULL : val, depth) != 0) { fprintf(stderr, "%s:%d: bad option `%s`\n", path, no, b); rc = -1; }
    }
    fclose(f);
    return rc;
}

static void config_print(const config* c, int hauteur_image)
{
    printf("image %dx%d spp %d bounces %d ao %d ao_intensity %g vfov %g focus %g aperture %g,%g seed %llu gpus %d\n", c->largeur_image, hauteur_image,
           c->nbRayonParPixel, c->nbRebondMax, c->useAO, c->AO_intensity, c->vfov, c->focus_distance, c->ouverture[0], c->ouverture[1], c->seed, c->n_gpus);
    printf("origin %g,%g,%g target %g,%g,%g up %g,%g,%g\n", c->origin[0], c->origin[1], c->origin[2], c->target[0], c->target[1], c->target[2], c->up[0], c->up[1], c->up[2]);
    printf("obj %s mtl %s sky %s out %s\n", c->obj_file ? c->obj_file : "-", c->mtl_file ? c->mtl_file : "-", c->sky_file ? c->sky_file : "-", c->out_file);
    printf("spheres %d\n", c->nbSpheres);
    for (int k = 0; k < c->nbSpheres; k++) {
        const trt_sphere* s = &c->spheres[k];
        printf("  sphere %g %g %g %g  %g %g %g  %g %g %g  %g %g %g %g\n", s->center.e[0], s->center.e[1], s->center.e[2], s->radius,
               s->mat.diffuseColor.e[0], s->mat.diffuseColor.e[1], s->mat.diffuseColor.e[2], s->mat.emissionColor.e[0], s->mat.emissionColor.e[1],
               s->mat.emissionColor.e[2], s->mat.emissionStrength, s->mat.reflectionStrength, s->mat.alpha, s->mat.materialIndex);
    }
    printf("overrides %d\n", c->n_overrides);
    for (int k = 0; k < c->n_overrides; k++) {
        const trt_mat_override* o = &c->overrides[k];
        printf("  override %d mask %u emission %g/%g/%g strength %g reflection %g alpha %g index %g\n", o->material_ordinal, o->mask, o->emissionColor[0],
               o->emissionColor[1], o->emissionColor[2], o->emissionStrength, o->reflectionStrength, o->alpha, o->materialIndex);
    }
}

int main(int argc, char** argv)
{
    /* ---- constants (main.c:293-347), README values, then the options ------------------------------------ */
    static config cfg;
    config_defaults(&cfg);
    for (int k = 1; k < argc; k++) {
        const char* a = argv[k];
        if (strncmp(a, "--", 2) != 0) { fprintf(stderr, "usage: see the header of apps/main_b200.c (bad option `%s`)\n", a); return 2; }
        const char* v = NULL;
        if (!is_flag(a + 2)) { if (k + 1 >= argc) { fprintf(stderr, "option `%s` needs a value\n", a); return 2; } v = argv[++k]; }
        if (config_set(&cfg, a + 2, v, 0) != 0) { fprintf(stderr, "usage: see the header of apps/main_b200.c (bad option `%s`)\n", a); return 2; }
    }
    const double ratio = cfg.ratio, vfov = cfg.vfov, focus_distance = cfg.focus_distance, AO_intensity = cfg.AO_intensity;
    const int largeur_image = cfg.largeur_image, nbRayonParPixel = cfg.nbRayonParPixel, nbRebondMax = cfg.nbRebondMax, useAO = cfg.useAO;
    const double *origin = cfg.origin, *target = cfg.target, *up = cfg.up, *ouverture = cfg.ouverture;
    const char *obj_file = cfg.obj_file, *mtl_file = cfg.mtl_file, *out_file = cfg.out_file, *sky_file = cfg.sky_file;
    const unsigned long long seed = cfg.seed; const int n_gpus = cfg.n_gpus, truncate_like_main_c = cfg.truncate_like_main_c, tile_split = cfg.tile_split;
    int hauteur_image = (int)(largeur_image / ratio);

    if (cfg.dry_run) { config_print(&cfg, hauteur_image); return 0; }
    trt_sphere* sphere_list = cfg.nbSpheres ? cfg.spheres : NULL; const int nbSpheres = cfg.nbSpheres;

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
    /* main.c:385-400 mallocs these three; page-locked memory from the library instead lets the GPUs write them directly over PCIe
       (with N GPUs N bands at once).  --pageable-output keeps main.c's plain allocation: same image, staged copies. */
    trt_vec3 *canva = NULL, *albedo_tab = NULL, *normal_tab = NULL;
    int pinned_out = !cfg.pageable_output && trt_host_alloc(npix * sizeof(trt_vec3), (void**)&canva) == TRT_OK &&
                     trt_host_alloc(npix * sizeof(trt_vec3), (void**)&albedo_tab) == TRT_OK && trt_host_alloc(npix * sizeof(trt_vec3), (void**)&normal_tab) == TRT_OK;
    if (!pinned_out) {
        trt_host_free(canva); trt_host_free(albedo_tab); trt_host_free(normal_tab);
        canva = (trt_vec3*)calloc(npix, sizeof(trt_vec3));
        albedo_tab = (trt_vec3*)calloc(npix, sizeof(trt_vec3));
        normal_tab = (trt_vec3*)calloc(npix, sizeof(trt_vec3));
    } else {
        memset(canva, 0, npix * sizeof(trt_vec3)); memset(albedo_tab, 0, npix * sizeof(trt_vec3)); memset(normal_tab, 0, npix * sizeof(trt_vec3));
    }

    /* ---- the replacement for main.c:402-453 -------------------------------------------------------------- */
    trt_ctx* ctx = NULL;
    void* comm = NULL;
    if (cfg.n_ranks > 1) {
        /* one process per GPU: this process drives one device; the 128-byte NCCL id travels through a file */
        if (!cfg.id_file || cfg.rank < 0 || cfg.rank >= cfg.n_ranks) { fprintf(stderr, "--nranks needs --rank 0..N-1 and --id-file\n"); return 2; }
        int dev = cfg.device >= 0 ? cfg.device : cfg.rank;
        if (trt_init(&dev, 1, &ctx) != TRT_OK) { fprintf(stderr, "trt_init: %s\n", trt_last_error(NULL)); return 1; }
        unsigned char id[TRT_NCCL_UNIQUE_ID_BYTES];
        if (cfg.rank == 0) {
            if (trt_nccl_unique_id(id) != TRT_OK) { fprintf(stderr, "trt_nccl_unique_id: %s\n", trt_last_error(NULL)); return 1; }
            char tmp[1024]; snprintf(tmp, sizeof tmp, "%s.tmp", cfg.id_file);
            FILE* f = fopen(tmp, "wb");
            if (!f || fwrite(id, 1, sizeof id, f) != sizeof id) { fprintf(stderr, "cannot write %s\n", tmp); return 1; }
            fclose(f);
            if (rename(tmp, cfg.id_file) != 0) { fprintf(stderr, "cannot rename %s\n", tmp); return 1; }
        } else {
            size_t got = 0;
            for (int tries = 0; tries < 6000 && got != sizeof id; tries++) {            /* up to 60 s */
                FILE* f = fopen(cfg.id_file, "rb");
                if (f) { got = fread(id, 1, sizeof id, f); fclose(f); }
                if (got != sizeof id) { struct timespec ts = {0, 10000000}; nanosleep(&ts, NULL); }
            }
            if (got != sizeof id) { fprintf(stderr, "rank %d: no NCCL id in %s\n", cfg.rank, cfg.id_file); return 1; }
        }
        if (trt_nccl_comm_init(ctx, cfg.n_ranks, cfg.rank, id, &comm) != TRT_OK) { fprintf(stderr, "trt_nccl_comm_init: %s\n", trt_last_error(ctx)); return 1; }
    } else if (trt_init(NULL, n_gpus, &ctx) != TRT_OK) { fprintf(stderr, "trt_init: %s\n", trt_last_error(NULL)); return 1; }
    if (trt_scene_set(ctx, sphere_list, nbSpheres, mesh.triangles, mesh.nbTriangles, mat_list, tex_width, tex_height,
                      mesh.nbMaterials, mesh.quelMatPourTri, cfg.overrides, cfg.n_overrides) != TRT_OK) {
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
    if (cfg.n_ranks > 1) {
        if (trt_render_distributed(ctx, &p, comm, cfg.rank, cfg.n_ranks, canva, albedo_tab, normal_tab, &st) != TRT_OK) {
            fprintf(stderr, "trt_render_distributed: %s\n", trt_last_error(ctx)); return 1;
        }
    } else if (trt_render(ctx, &p, canva, albedo_tab, normal_tab, &st) != TRT_OK) { fprintf(stderr, "trt_render: %s\n", trt_last_error(ctx)); return 1; }
    t_render = now_s() - t_render;
    if (cfg.n_ranks > 1 && cfg.rank != 0) {                   /* the image is on rank 0 */
        printf("rank %d of %d: %llu paths, %llu rays, trace %.2f ms\n", cfg.rank, cfg.n_ranks, (unsigned long long)st.paths, (unsigned long long)st.rays, st.ms_trace);
        trt_nccl_comm_destroy(comm); trt_destroy(ctx);
        return 0;
    }

    /* useDenoiser (main.c:455) stays an external post-pass on canva / albedo_tab / normal_tab */
    if (trt_write_ppm(out_file, canva, largeur_image, hauteur_image) != 0) { fprintf(stderr, "cannot write %s\n", out_file); return 1; }

    printf("%dx%d, %d spp, %d bounces, AO %s: %llu paths, %llu rays, trace %.2f ms, render call %.2f ms (%.1f Mrays/s), total %.3f s -> %s\n",
           largeur_image, hauteur_image, nbRayonParPixel, nbRebondMax, useAO ? "on" : "off",
           (unsigned long long)st.paths, (unsigned long long)st.rays, st.ms_trace, t_render * 1e3,
           (double)st.rays / t_render / 1e6, now_s() - t_start, out_file);

    if (comm) trt_nccl_comm_destroy(comm);
    trt_destroy(ctx);
    if (pinned_out) { trt_host_free(canva); trt_host_free(albedo_tab); trt_host_free(normal_tab); }
    else { free(canva); free(albedo_tab); free(normal_tab); }
    free(mat_list); free(sky_mat_list);
    trt_free_mesh(&mesh);
    return 0;
}
