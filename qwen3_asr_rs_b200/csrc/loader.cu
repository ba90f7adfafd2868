#include "internal.h"
namespace asrb {
void model_load_dir(Ctx*, const char*, Model**) { throw Error(ASRB_ERR_IO, "asrb_model_load: not implemented yet"); }
}
