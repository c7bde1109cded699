/* oracle/ref_atmosphere.cu -- TEST INFRASTRUCTURE (oracle side, never the product).
 *
 * Drives the reference's OWN Bruneton sky precompute so both render kernels (the unmodified reference
 * cubin and libvpt_b200) can be handed the same `AtmosphereParameters` -- scalars plus the four
 * look-up textures -- for environment_type == 0 parity tests.  The reference sources are compiled from
 * /root/reference at build time (oracle/Makefile: source/atmosphere/atmosphere.cpp and
 * atmosphere_kernels.cu); nothing is copied into the repo.
 *
 *   vptref_atmosphere_init : source/main.cpp:1468-1472 (`earth_atmosphere.init()`), with the GUI's default
 *                            switches of main.cpp:1433-1436 passed in by the caller.  The reference loads
 *                            "atmosphere_kernels.ptx" from the working directory (atmosphere.cpp:1190), so the
 *                            call runs with the cwd switched to `module_dir` (oracle/_ref/atmo).
 *
 * The reference's util/fileIO.cpp (OpenEXR/stb) and util/logger.cpp are not built; the five functions of
 * theirs that atmosphere.cpp links against are stubbed below: loaders report "no cached textures" (which makes
 * init() precompute, atmosphere.cpp:1182-1185), savers are no-ops.
 */
#include <cstdio>
#include <cstring>
#include <string>
#include <new>
#include <cstdlib>
#include <unistd.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include "helper_math.h"
#include "atmosphere/atmosphere.h"

bool load_texture_exr(float4**, std::string, int&, int&, bool) { return false; }
bool load_texture_exr_gpu(float4**, std::string, int&, int&, bool) { return false; }
bool save_texture_exr(float4*, std::string, const int, const int, bool) { return true; }
bool save_texture_png(float4*, std::string, const int, const int) { return true; }
void log(const char*, unsigned int) {}
void log(std::string, unsigned int) {}

static atmosphere* g_atmo = nullptr;

extern "C" int vptref_atmosphere_init(const char* module_dir, int use_constant_solar_spectrum, int use_ozone,
                                      int luminance_mode, int do_white_balance, float exposure, void* out_params)
{
    static_assert(sizeof(AtmosphereParameters) == 464, "AtmosphereParameters layout");
    cudaFree(0);                                     // make the runtime's primary context current for cuModuleLoad
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) return -1;
    if (chdir(module_dir) != 0) { fprintf(stderr, "[vptref] atmosphere: cannot enter %s\n", module_dir); return -2; }
    // A fresh model per call (init() appends to its spectra); the old one is kept alive for its textures.  The reference's
    // object has static storage (main.cpp: `atmosphere earth_atmosphere;`) and relies on that zero-initialisation:
    // copy_*_texture() destroys "the previous texture" whenever the handle member is non-zero (atmosphere.cpp:506).
    // Give the object the same zeroed storage here, otherwise a stale heap word is destroyed as a texture handle.
    void* storage = calloc(1, sizeof(atmosphere));
    g_atmo = new (storage) atmosphere();
    {
        // The reference's precompute kernels index their scattering buffers with unclamped texel coordinates
        // (atmosphere_kernels.cu:378-392: u = 1 maps to texel `size`, one row / slice past the end, up to ~0.5 MB beyond
        // the allocation), which faults or not depending on what the driver mapped next to the buffer.  The buffer
        // pointers are public members, so the harness re-homes the nine buffers into one zeroed slab with 2 MB of
        // slack on both sides of each: the stray reads then land in mapped zeros, as they do in a lucky run.
        AtmosphereParameters& P = g_atmo->atmosphere_parameters;
        const size_t small = (size_t)TRANSMITTANCE_TEXTURE_WIDTH * TRANSMITTANCE_TEXTURE_HEIGHT * sizeof(float4);
        const size_t irr = (size_t)IRRADIANCE_TEXTURE_WIDTH * IRRADIANCE_TEXTURE_HEIGHT * sizeof(float4);
        const size_t big = (size_t)SCATTERING_TEXTURE_WIDTH * SCATTERING_TEXTURE_HEIGHT * SCATTERING_TEXTURE_DEPTH * sizeof(float4);
        float4** slot[9] = { &P.transmittance_buffer, &P.delta_irradience_buffer, &P.irradiance_buffer, &P.delta_rayleigh_scattering_buffer,
                             &P.delta_mie_scattering_buffer, &P.scattering_buffer, &P.optional_mie_single_scattering_buffer,
                             &P.delta_scattering_density_buffer, &P.delta_multiple_scattering_buffer };
        const size_t bytes[9] = { small, irr, irr, big, big, big, big, big, big };
        const size_t slack = (size_t)2 << 20;
        size_t total = slack;
        for (int i = 0; i < 9; ++i) total += bytes[i] + slack;
        char* slab = nullptr;
        if (cudaMalloc(&slab, total) != cudaSuccess || cudaMemset(slab, 0, total) != cudaSuccess) { chdir(cwd); return -5; }
        size_t off = slack;
        for (int i = 0; i < 9; ++i) { cudaFree(*slot[i]); *slot[i] = reinterpret_cast<float4*>(slab + off); off += bytes[i] + slack; }
    }
    g_atmo->m_use_constant_solar_spectrum = use_constant_solar_spectrum != 0;
    g_atmo->m_use_ozone = use_ozone != 0;
    g_atmo->m_use_luminance = luminance_mode == 1 ? APPROXIMATE : luminance_mode == 2 ? PRECOMPUTED : NONE;
    g_atmo->m_do_white_balance = do_white_balance != 0;
    g_atmo->m_exposure = exposure;
    g_atmo->texture_folder = "";                     // no cached EXRs: always precompute
    atmosphere_error_t err = g_atmo->init();
    if (chdir(cwd) != 0) return -3;
    if (err != ATMO_NO_ERR) return -10 - (int)err;
    if (cudaDeviceSynchronize() != cudaSuccess) return -4;
    g_atmo->update_model();                          // main.cpp:1733-1738 (exposure / white balance refresh)
    memcpy(out_params, &g_atmo->atmosphere_parameters, sizeof(AtmosphereParameters));
    return 0;
}
