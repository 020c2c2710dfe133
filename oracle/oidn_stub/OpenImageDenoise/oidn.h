/* TEST INFRASTRUCTURE — not shipped, never linked into the product library.
 *
 * Minimal stand-in for <OpenImageDenoise/oidn.h> so that the reference's main.c
 * (which includes it at main.c:9 and uses it only inside denoiser(), denoiser.h:31-91)
 * can be compiled as the parity oracle on a machine without libOpenImageDenoise.
 * The denoiser is a post-pass outside the hot path (main.c:455, useDenoiser=false at
 * main.c:314); every entry point below is an inert placeholder that is never reached.
 */
#ifndef TRT_ORACLE_OIDN_STUB_H
#define TRT_ORACLE_OIDN_STUB_H

#include <stddef.h>

typedef struct trt_stub_oidn_device* OIDNDevice;
typedef struct trt_stub_oidn_buffer* OIDNBuffer;
typedef struct trt_stub_oidn_filter* OIDNFilter;

typedef enum { OIDN_DEVICE_TYPE_DEFAULT = 0 } OIDNDeviceType;
typedef enum { OIDN_FORMAT_UNDEFINED = 0, OIDN_FORMAT_FLOAT3 = 3 } OIDNFormat;
typedef enum { OIDN_ERROR_NONE = 0, OIDN_ERROR_UNKNOWN = 1 } OIDNError;

static inline OIDNDevice oidnNewDevice(OIDNDeviceType t) { (void)t; return NULL; }
static inline void oidnCommitDevice(OIDNDevice d) { (void)d; }
static inline OIDNBuffer oidnNewBuffer(OIDNDevice d, size_t n) { (void)d; (void)n; return NULL; }
static inline void* oidnGetBufferData(OIDNBuffer b) { (void)b; return NULL; }
static inline OIDNFilter oidnNewFilter(OIDNDevice d, const char* kind) { (void)d; (void)kind; return NULL; }
static inline void oidnSetFilterImage(OIDNFilter f, const char* name, OIDNBuffer b, OIDNFormat fmt,
                                      size_t w, size_t h, size_t off, size_t pstride, size_t rstride)
{ (void)f; (void)name; (void)b; (void)fmt; (void)w; (void)h; (void)off; (void)pstride; (void)rstride; }
static inline void oidnCommitFilter(OIDNFilter f) { (void)f; }
static inline void oidnExecuteFilter(OIDNFilter f) { (void)f; }
static inline OIDNError oidnGetDeviceError(OIDNDevice d, const char** msg)
{ (void)d; if (msg) *msg = "oidn stub"; return OIDN_ERROR_UNKNOWN; }
static inline void oidnReleaseBuffer(OIDNBuffer b) { (void)b; }
static inline void oidnReleaseFilter(OIDNFilter f) { (void)f; }
static inline void oidnReleaseDevice(OIDNDevice d) { (void)d; }

#endif
