// The reference's RendererTypes module (Sources/RendererTypes/include/BridgingTypes.h) on Linux.
#include "../../../../include/gsm/gsm_types.h"
