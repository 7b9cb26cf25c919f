// emu_b200 — placeholders for sub-models that are not wired yet (replaced file by file as they land).
#include "common.cuh"
#include "engine.h"

namespace emu {
struct UNetModel {};
struct VaeModel {};
struct CFormerModel {};
int unet_load_tensor(EmuEngine* e, const std::string&, const bf16*, const int64_t*, int, cudaStream_t) {
  return e->fail(EMU_ERR_UNSUPPORTED, "UNet not built in this library yet");
}
int vae_load_tensor(EmuEngine* e, const std::string&, const bf16*, const int64_t*, int, cudaStream_t) {
  return e->fail(EMU_ERR_UNSUPPORTED, "VAE not built in this library yet");
}
int cformer_load_tensor(EmuEngine* e, const std::string&, const bf16*, const int64_t*, int, cudaStream_t) {
  return e->fail(EMU_ERR_UNSUPPORTED, "Causal-Former not built in this library yet");
}
void unet_destroy(UNetModel* m) { delete m; }
void vae_destroy(VaeModel* m) { delete m; }
void cformer_destroy(CFormerModel* m) { delete m; }
}  // namespace emu

extern "C" int emu_cformer_forward(EmuEngine* e, const void*, int, int, void*, emu_stream_t) {
  return e ? e->fail(EMU_ERR_UNSUPPORTED, "Causal-Former not built yet") : EMU_ERR_INVALID;
}
extern "C" int emu_unet_configure(EmuEngine* e, const EmuUNetConfig*) {
  return e ? e->fail(EMU_ERR_UNSUPPORTED, "UNet not built yet") : EMU_ERR_INVALID;
}
extern "C" int emu_unet_forward(EmuEngine* e, const void*, float, const void*, int, const void*, const int32_t*, int, int,
                                int, void*, emu_stream_t) {
  return e ? e->fail(EMU_ERR_UNSUPPORTED, "UNet not built yet") : EMU_ERR_INVALID;
}
extern "C" int emu_denoise_step(EmuEngine* e, float*, float, float, float, float, const void*, int, const void*,
                                const int32_t*, int, int, int, emu_stream_t) {
  return e ? e->fail(EMU_ERR_UNSUPPORTED, "UNet not built yet") : EMU_ERR_INVALID;
}
extern "C" int emu_vae_configure(EmuEngine* e, const EmuVAEConfig*) {
  return e ? e->fail(EMU_ERR_UNSUPPORTED, "VAE not built yet") : EMU_ERR_INVALID;
}
extern "C" int emu_vae_decode(EmuEngine* e, const void*, int, int, int, float*, emu_stream_t) {
  return e ? e->fail(EMU_ERR_UNSUPPORTED, "VAE not built yet") : EMU_ERR_INVALID;
}
