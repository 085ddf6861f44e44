// stubs.cu -- entry points not built yet in this round; they fail loudly (B2S_ERR_UNSUPPORTED), never fall back.
#include "b2s_internal.h"
extern "C" int b2s_slab_config(b2s_ctx* ctx,int,int,uint32_t,uint32_t){ return b2s_fail(ctx,B2S_ERR_UNSUPPORTED,"slab decomposition not built"); }
extern "C" int b2s_slab_pack_halo(b2s_ctx* ctx,int,void**,uint32_t*){ return b2s_fail(ctx,B2S_ERR_UNSUPPORTED,"slab decomposition not built"); }
extern "C" int b2s_slab_unpack_halo(b2s_ctx* ctx,int,const void*,uint32_t){ return b2s_fail(ctx,B2S_ERR_UNSUPPORTED,"slab decomposition not built"); }
extern "C" int b2s_slab_pack_migrants(b2s_ctx* ctx,int,void**,uint32_t*){ return b2s_fail(ctx,B2S_ERR_UNSUPPORTED,"slab decomposition not built"); }
extern "C" int b2s_slab_unpack_migrants(b2s_ctx* ctx,int,const void*,uint32_t){ return b2s_fail(ctx,B2S_ERR_UNSUPPORTED,"slab decomposition not built"); }
extern "C" int b2s_slab_dt_buffer(b2s_ctx* ctx,void**){ return b2s_fail(ctx,B2S_ERR_UNSUPPORTED,"slab decomposition not built"); }
extern "C" int b2s_slab_apply_dt(b2s_ctx* ctx){ return b2s_fail(ctx,B2S_ERR_UNSUPPORTED,"slab decomposition not built"); }
