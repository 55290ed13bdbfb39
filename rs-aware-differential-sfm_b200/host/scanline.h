// scanline.h -- same header name as the reference's src/scanline.h: the declarations live in rsdsfm_host.h.
#pragma once
#include "rsdsfm_host.h"
