/* oracle/pt_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * C interface of the CPU restatement of the reference's hot path (see
 * pt_oracle.cpp).  It consumes the same plain-old-data descriptors as the
 * product's C ABI (include/b200pt.h) so tests can feed identical inputs to
 * both.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#ifndef B200PT_ORACLE_H
#define B200PT_ORACLE_H

#include <stdint.h>

#include "../include/b200pt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_scene oracle_scene;

oracle_scene *oracle_scene_create(const b200pt_scene_desc *desc);
void oracle_scene_destroy(oracle_scene *s);

/* Scene::Intersect / IntersectP on a batch of rays through a binary BVH
 * traversed exactly like accelerators/bvh.cpp:662-738. */
int oracle_trace_closest(const oracle_scene *s, const b200pt_ray *rays, b200pt_hit *hits, int64_t n);
int oracle_trace_any(const oracle_scene *s, const b200pt_ray *rays, uint8_t *occluded, int64_t n);
/* Same answers by testing every triangle in index order (no BVH). */
int oracle_trace_closest_brute(const oracle_scene *s, const b200pt_ray *rays, b200pt_hit *hits,
                               int64_t n);

/* SamplerIntegrator::Render over the given tiles (NULL = all), accumulating
 * into film_xyzw ([h][w][4] over the cropped bounds: X,Y,Z,weight; the caller
 * zero-initialises it).  n_threads host threads share the tiles. */
int oracle_render(const oracle_scene *s, const b200pt_camera_desc *camera,
                  const b200pt_film_desc *film, const b200pt_sampler_desc *sampler,
                  const b200pt_integrator_desc *integrator, const int32_t *tiles, int64_t n_tiles,
                  int n_threads, float *film_xyzw, b200pt_stats *stats);
/* Film::WriteImage pixel pipeline (film.cpp:174-203). */
int oracle_film_rgb(const b200pt_film_desc *film, const float *film_xyzw, float *rgb);

int oracle_sobol(const b200pt_sampler_desc *sampler, int32_t px, int32_t py, int64_t sample,
                 int32_t dim0, int32_t n_dims, float *out);
/* HaltonSampler::radicalInversePermutations for the first n_bases primes (lowdiscrepancy.cpp:2490-2504);
 * returns the number of entries written (PrimeSums[n_bases]); out may be NULL to query the size. */
int64_t oracle_halton_permutations(int32_t n_bases, uint16_t *out);
int oracle_camera_rays(const b200pt_camera_desc *camera, const b200pt_sampler_desc *sampler,
                       int32_t px, int32_t py, int32_t n_samples, b200pt_ray *out);
/* Per-sample radiance of one pixel after the guards of integrator.cpp:294-315. */
int oracle_pixel_samples(const oracle_scene *s, const b200pt_camera_desc *camera,
                         const b200pt_film_desc *film, const b200pt_sampler_desc *sampler,
                         const b200pt_integrator_desc *integrator, int32_t px, int32_t py,
                         float *out_rgb);

/* Sphere::Sample(ref, u, pdf) and Sphere::Pdf(ref, wi) (sphere.cpp:232-304) for one sphere of a descriptor:
 * out = {p.xyz, n.xyz, pError.xyz, pdf}.  Used by tests/host_preflight.cpp to check the device routines. */
void oracle_sphere_sample(const b200pt_sphere *sphere, const float ref_p[3], const float ref_perr[3], const float ref_n[3],
                          const float u[2], float out[10]);
float oracle_sphere_pdf(const b200pt_sphere *sphere, const float ref_p[3], const float ref_perr[3], const float ref_n[3],
                        const float wi[3]);

/* liboracle_spectral.so only (ORACLE_NSPEC == 60): the reference's spectral data.  oracle_spectral_register maps a
 * descriptor RGB triple to the 60-bin SampledSpectrum the reference derives for it (Spectrum::FromRGB for "rgb"
 * parameters, a constant for float defaults, FromSampled for the metal's measured eta / k); oracle_spectral_set_cie
 * passes SampledSpectrum::X, Y, Z.  Both come from fixtures dumped by the spectral probe.  Return the bin count. */
int oracle_spectral_register(const float rgb[3], const float *spectrum);
int oracle_spectral_set_cie(const float *X, const float *Y, const float *Z);
/* VolPathIntegrator (integrators/volpath.cpp) with at most one homogeneous medium around the whole scene; groundwork,
 * the product ABI has no media yet */
void oracle_set_volpath(int enabled, int has_medium, const float sigma_a[3], const float sigma_s[3], float g);
/* groundwork: media bounded by null-material spheres (`Material ""` under `MediumInterface "cloud" ""`) */
void oracle_set_medium_boundaries(int n, const int *sphere_index, const float *sigma_a, const float *sigma_s, const float *g);
int oracle_spectrum_samples(void);

/* The host libm's sinf/cosf (what the reference calls through std::sin/cos). */
float oracle_libm_sinf(float x);
float oracle_libm_cosf(float x);

#ifdef __cplusplus
}
#endif
#endif
