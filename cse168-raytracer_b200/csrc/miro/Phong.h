// Phong.h -- forwarding header (reference Phong.h): class Phong lives in Material.h of the host API layer.
#ifndef MIROHOST_PHONG_H
#define MIROHOST_PHONG_H
#include "Material.h"
#endif
