// minimal.h -- same header name as the reference's src/minimal.h: the declarations live in rsdsfm_host.h.
#pragma once
#include "rsdsfm_host.h"
