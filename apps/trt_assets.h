/* trt_assets.h — host-side asset I/O for main_b200.c: OBJ/MTL meshes and P3 PPM textures, in plain C.
 *
 * Same file formats, conventions and output arrays as the reference's loaders, re-implemented (not copied):
 *   trt_load_mesh      ~ list_of_mesh         (mesh.h:96-218)   + tex_path_from_mtl (rtutility.h:233-290)
 *   trt_load_textures  ~ create_mat_list_mtl  (texture.h:175-354)
 *   trt_init_camera    ~ init_camera          (camera.h:17-40)
 *   trt_write_ppm      ~ the P3 writer        (main.c:457-465)
 * so the arrays they return can be handed to trt_scene_set exactly like main.c's own.
 *
 * Conventions kept from the reference:
 *   - faces are triangles written `f v/vt/vn v/vt/vn v/vt/vn`; indices are 1-based;
 *   - a triangle's material ordinal is the number of `usemtl` lines seen before it, minus one;
 *   - the texture of material `name` is the first `map_Kd` after `newmtl name` in the MTL; the MTL names a
 *     .png, the loader reads `<stem>.ppm` (colour) and `<stem>_alpha.ppm` (alpha = first channel), P3 only;
 *   - texel row 0 is the BOTTOM row of the image (texture.h:227), values are sample/maxVal;
 *   - material fields the reference leaves unset (texture.h:244-246) are 0 here.
 * What is new: files are read in one go and parsed by hand (the reference's fscanf loop dominates end-to-end
 * time once the render takes milliseconds), and every failure is reported instead of crashing later.
 */
#ifndef TRT_ASSETS_H
#define TRT_ASSETS_H

#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/tipe_rt.h"

typedef struct trt_mesh {
    trt_triangle* triangles;   int nbTriangles;
    int*          quelMatPourTri;
    char**        texture_paths;  int nbMaterials;     /* as named by the MTL (usually .png) */
} trt_mesh;

static char* trt__slurp(const char* path, size_t* len)
{
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    char* buf = (char*)malloc((size_t)n + 1);
    if (!buf) { fclose(f); return NULL; }
    size_t got = fread(buf, 1, (size_t)n, f);
    fclose(f);
    buf[got] = '\0';
    if (len) *len = got;
    return buf;
}

static void trt__rstrip(char* s)
{
    size_t n = strlen(s);
    while (n > 0 && (s[n - 1] == '\n' || s[n - 1] == '\r' || s[n - 1] == ' ' || s[n - 1] == '\t')) s[--n] = '\0';
}

/* First map_Kd after `newmtl <name>`, resolved against the MTL's directory, leading "./" dropped. */
static char* trt__texture_of(const char* mtl_text, const char* mtl_path, const char* name)
{
    const char* p = mtl_text;
    while ((p = strstr(p, "newmtl ")) != NULL) {
        if (p != mtl_text && p[-1] != '\n') { p += 7; continue; }
        char line[512];
        size_t k = 0;
        const char* q = p + 7;
        while (*q && *q != '\n' && k < sizeof line - 1) line[k++] = *q++;
        line[k] = '\0';
        trt__rstrip(line);
        if (strcmp(line, name) == 0) {
            const char* m = q;
            while ((m = strstr(m, "map_Kd ")) != NULL) {
                if (m[-1] != '\n') { m += 7; continue; }
                k = 0; m += 7;
                while (*m && *m != '\n' && k < sizeof line - 1) line[k++] = *m++;
                line[k] = '\0';
                trt__rstrip(line);
                const char* rel = strncmp(line, "./", 2) == 0 ? line + 2 : line;
                const char* slash = strrchr(mtl_path, '/');
                size_t dir = slash ? (size_t)(slash - mtl_path + 1) : 0;
                char* out = (char*)malloc(dir + strlen(rel) + 1);
                memcpy(out, mtl_path, dir);
                strcpy(out + dir, rel);
                return out;
            }
            return NULL;
        }
        p = q;
    }
    return NULL;
}

static void trt_free_mesh(trt_mesh* m)
{
    if (!m) return;
    free(m->triangles); free(m->quelMatPourTri);
    for (int k = 0; k < m->nbMaterials; k++) free(m->texture_paths[k]);
    free(m->texture_paths);
    memset(m, 0, sizeof *m);
}

/* Returns 0 on success; on failure err (>= 256 bytes) holds the reason. */
static int trt_load_mesh(const char* obj_path, const char* mtl_path, trt_mesh* out, char* err)
{
    memset(out, 0, sizeof *out);
    size_t len = 0;
    char* obj = trt__slurp(obj_path, &len);
    if (!obj) { snprintf(err, 256, "cannot open mesh file %s", obj_path); return -1; }
    char* mtl = trt__slurp(mtl_path, NULL);
    if (!mtl) { snprintf(err, 256, "cannot open material file %s", mtl_path); free(obj); return -1; }

    int nv = 0, nvt = 0, nf = 0, nm = 0;
    for (char* p = obj; *p;) {
        if (p[0] == 'v' && p[1] == ' ') nv++;
        else if (p[0] == 'v' && p[1] == 't') nvt++;
        else if (p[0] == 'f') nf++;
        else if (strncmp(p, "usemtl", 6) == 0) nm++;
        char* e = strchr(p, '\n');
        if (!e) break;
        p = e + 1;
    }
    double* v = (double*)malloc(sizeof(double) * 3 * (size_t)(nv > 0 ? nv : 1));
    double* vt = (double*)malloc(sizeof(double) * 2 * (size_t)(nvt > 0 ? nvt : 1));
    out->triangles = (trt_triangle*)calloc((size_t)(nf > 0 ? nf : 1), sizeof(trt_triangle));
    out->quelMatPourTri = (int*)malloc(sizeof(int) * (size_t)(nf > 0 ? nf : 1));
    out->texture_paths = (char**)calloc((size_t)(nm > 0 ? nm : 1), sizeof(char*));
    int iv = 0, ivt = 0, itri = 0, imat = -1, rc = 0;
    for (char* p = obj; *p && rc == 0;) {
        char* e = strchr(p, '\n');
        if (e) *e = '\0';
        if (p[0] == 'v' && p[1] == ' ') {
            char* q = p + 2;
            for (int a = 0; a < 3; a++) v[3 * iv + a] = strtod(q, &q);
            iv++;
        } else if (p[0] == 'v' && p[1] == 't') {
            char* q = p + 2;
            vt[2 * ivt] = strtod(q, &q); vt[2 * ivt + 1] = strtod(q, &q);
            ivt++;
        } else if (p[0] == 'f') {
            int vi[3], ti[3];
            if (sscanf(p, "f %d/%d/%*d %d/%d/%*d %d/%d/%*d", &vi[0], &ti[0], &vi[1], &ti[1], &vi[2], &ti[2]) != 6) {
                snprintf(err, 256, "%s: face %d is not a `v/vt/vn` triangle", obj_path, itri); rc = -1;
            } else {
                trt_triangle* t = &out->triangles[itri];
                for (int c = 0; c < 3; c++) {
                    if (vi[c] < 1 || vi[c] > nv || ti[c] < 1 || ti[c] > nvt) { snprintf(err, 256, "%s: face %d index out of range", obj_path, itri); rc = -1; break; }
                    trt_vec3* dst = c == 0 ? &t->A : (c == 1 ? &t->B : &t->C);
                    trt_uv* uv = c == 0 ? &t->uvA : (c == 1 ? &t->uvB : &t->uvC);
                    memcpy(dst->e, v + 3 * (vi[c] - 1), 24);
                    uv->u = vt[2 * (ti[c] - 1)]; uv->v = vt[2 * (ti[c] - 1) + 1];
                }
                t->mat.diffuseColor.e[0] = 0.784; t->mat.diffuseColor.e[1] = 0.965; t->mat.diffuseColor.e[2] = 1.0;   /* SKY, mesh.h:205 */
                if (imat < 0 && rc == 0) { snprintf(err, 256, "%s: face before any usemtl", obj_path); rc = -1; }
                out->quelMatPourTri[itri++] = imat;
            }
        } else if (strncmp(p, "usemtl ", 7) == 0) {
            char name[512];
            strncpy(name, p + 7, sizeof name - 1); name[sizeof name - 1] = '\0';
            trt__rstrip(name);
            imat++;
            out->texture_paths[imat] = trt__texture_of(mtl, mtl_path, name);
            if (!out->texture_paths[imat]) { snprintf(err, 256, "%s: no map_Kd for material `%s`", mtl_path, name); rc = -1; }
        }
        if (!e) break;
        p = e + 1;
    }
    out->nbTriangles = itri; out->nbMaterials = imat + 1;
    free(v); free(vt); free(obj); free(mtl);
    if (rc != 0) trt_free_mesh(out);
    return rc;
}

/* move_mesh (mesh.h:220-234) */
static void trt_move_mesh(trt_mesh* m, double x, double y, double z)
{
    for (int k = 0; k < m->nbTriangles; k++) {
        trt_vec3* p[3] = {&m->triangles[k].A, &m->triangles[k].B, &m->triangles[k].C};
        for (int c = 0; c < 3; c++) { p[c]->e[0] += x; p[c]->e[1] += y; p[c]->e[2] += z; }
    }
}

/* Hand parser for `P3 <w> <h> <max> <samples...>` (comments allowed). Returns malloc'd ints [h*w*3] in file order. */
static int* trt__read_p3(const char* path, int* w, int* h, int* maxval, char* err)
{
    size_t len = 0;
    char* txt = trt__slurp(path, &len);
    if (!txt) { snprintf(err, 256, "cannot open texture %s", path); return NULL; }
    const char* p = txt;
    if (!(p[0] == 'P' && p[1] == '3')) { snprintf(err, 256, "%s is not a P3 PPM", path); free(txt); return NULL; }
    p += 2;
    int hdr[3], got = 0;
    long long count = 0, total = 0;
    int* data = NULL;
    while (*p) {
        while (*p && isspace((unsigned char)*p)) p++;
        if (*p == '#') { while (*p && *p != '\n') p++; continue; }
        if (!*p) break;
        if (!isdigit((unsigned char)*p)) { snprintf(err, 256, "%s: unexpected character in PPM", path); free(txt); free(data); return NULL; }
        int val = 0;
        while (isdigit((unsigned char)*p)) val = val * 10 + (*p++ - '0');
        if (got < 3) {
            hdr[got++] = val;
            if (got == 3) {
                if (hdr[0] < 1 || hdr[1] < 1 || hdr[2] < 1) { snprintf(err, 256, "%s: bad PPM header", path); free(txt); return NULL; }
                total = (long long)hdr[0] * hdr[1] * 3;
                data = (int*)malloc(sizeof(int) * (size_t)total);
                if (!data) { snprintf(err, 256, "%s: out of memory", path); free(txt); return NULL; }
            }
        } else if (count < total) data[count++] = val;
    }
    free(txt);
    if (got < 3 || count < total) { snprintf(err, 256, "%s: truncated PPM (%lld of %lld samples)", path, count, total); free(data); return NULL; }
    *w = hdr[0]; *h = hdr[1]; *maxval = hdr[2];
    return data;
}

/* One slab of tex_h*tex_w materials per usemtl, like create_mat_list_mtl. Returns NULL and fills err on failure. */
static trt_material* trt_load_textures(char** texture_paths, int nbMaterials, int* tex_w, int* tex_h, char* err)
{
    trt_material* out = NULL;
    for (int k = 0; k < nbMaterials; k++) {
        const char* src = texture_paths[k];
        const char* ext = strstr(src, ".png");
        if (!ext) { snprintf(err, 256, "texture `%s` does not name a .png", src); free(out); return NULL; }
        size_t stem = (size_t)(ext - src);
        char* col = (char*)malloc(stem + 16); char* alp = (char*)malloc(stem + 16);
        memcpy(col, src, stem); strcpy(col + stem, ".ppm");
        memcpy(alp, src, stem); strcpy(alp + stem, "_alpha.ppm");
        int w = 0, h = 0, mx = 0, aw = 0, ah = 0, amx = 0;
        int* c = trt__read_p3(col, &w, &h, &mx, err);
        int* a = c ? trt__read_p3(alp, &aw, &ah, &amx, err) : NULL;
        free(col); free(alp);
        if (!c || !a) { free(c); free(a); free(out); return NULL; }
        if (aw != w || ah != h || (k > 0 && (w != *tex_w || h != *tex_h))) {
            snprintf(err, 256, "texture %d (`%s`): all textures and alpha maps must share one size", k, src);
            free(c); free(a); free(out); return NULL;
        }
        if (k == 0) {
            *tex_w = w; *tex_h = h;
            out = (trt_material*)calloc((size_t)nbMaterials * w * h, sizeof(trt_material));
            if (!out) { snprintf(err, 256, "out of memory for %d textures of %dx%d", nbMaterials, w, h); free(c); free(a); return NULL; }
        }
        trt_material* slab = out + (size_t)k * w * h;
        for (int row = 0; row < h; row++) {                       /* file row 0 is the top: store bottom-up */
            const int i = h - 1 - row;
            for (int j = 0; j < w; j++) {
                const int* px = c + ((size_t)row * w + j) * 3;
                trt_material* m = &slab[(size_t)i * w + j];
                m->diffuseColor.e[0] = (double)px[0] / mx; m->diffuseColor.e[1] = (double)px[1] / mx; m->diffuseColor.e[2] = (double)px[2] / mx;
                m->alpha = (double)a[((size_t)row * w + j) * 3] / amx;
                m->emissionStrength = 0.0;
            }
        }
        free(c); free(a);
    }
    return out;
}

/* Sky image, like create_mat_list (texture.h:145-173): sky_w*sky_h materials whose diffuseColor is the P3 texel / maxVal,
 * file row 0 (top) stored last, every other field 0.  For trt_sky_set. */
static trt_material* trt_load_sky(const char* path, int* sky_w, int* sky_h, char* err)
{
    int w = 0, h = 0, mx = 0;
    int* c = trt__read_p3(path, &w, &h, &mx, err);
    if (!c) return NULL;
    trt_material* out = (trt_material*)calloc((size_t)w * h, sizeof(trt_material));
    if (!out) { snprintf(err, 256, "out of memory for a %dx%d sky", w, h); free(c); return NULL; }
    for (int row = 0; row < h; row++) {
        const int i = h - 1 - row;
        for (int j = 0; j < w; j++) {
            const int* px = c + ((size_t)row * w + j) * 3;
            trt_material* m = &out[(size_t)i * w + j];
            m->diffuseColor.e[0] = (double)px[0] / mx; m->diffuseColor.e[1] = (double)px[1] / mx; m->diffuseColor.e[2] = (double)px[2] / mx;
        }
    }
    free(c);
    *sky_w = w; *sky_h = h;
    return out;
}

/* init_camera (camera.h:17-40), same operation order. */
static trt_camera trt_init_camera(const double origin[3], const double target[3], const double up[3], double vfov, double ratio)
{
    trt_camera cam;
    double theta = vfov * 3.1415926535897932385 / 180.0;
    double h = tan(theta / 2);
    double vh = 2.0 * h, vw = ratio * vh;
    double w[3], u[3], v[3], len;
    for (int a = 0; a < 3; a++) w[a] = origin[a] - target[a];
    len = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    for (int a = 0; a < 3; a++) w[a] = w[a] / len;
    u[0] = up[1] * w[2] - up[2] * w[1]; u[1] = up[2] * w[0] - up[0] * w[2]; u[2] = up[0] * w[1] - up[1] * w[0];
    len = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    for (int a = 0; a < 3; a++) u[a] = u[a] / len;
    v[0] = w[1] * u[2] - w[2] * u[1]; v[1] = w[2] * u[0] - w[0] * u[2]; v[2] = w[0] * u[1] - w[1] * u[0];
    for (int a = 0; a < 3; a++) {
        cam.origin.e[a] = origin[a];
        cam.horizontal.e[a] = u[a] * vw;
        cam.vertical.e[a] = v[a] * vh;
    }
    for (int a = 0; a < 3; a++)
        cam.coin_bas_gauche.e[a] = cam.origin.e[a] - (cam.horizontal.e[a] / 2 + (cam.vertical.e[a] / 2 + w[a]));
    return cam;
}

/* main.c:457-465 — P3, rows from the top (j = H-1) down. */
static int trt_write_ppm(const char* path, const trt_vec3* canva, int W, int H)
{
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    fprintf(f, "P3\n%d %d\n255\n", W, H);
    for (int j = H - 1; j >= 0; j--)
        for (int i = 0; i < W; i++) {
            const trt_vec3* c = &canva[(size_t)j * W + i];
            fprintf(f, "%d %d %d\n", (int)c->e[0], (int)c->e[1], (int)c->e[2]);
        }
    fclose(f);
    return 0;
}

#endif /* TRT_ASSETS_H */
