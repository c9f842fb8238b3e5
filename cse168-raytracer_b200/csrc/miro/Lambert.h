// Lambert.h -- forwarding header (reference Lambert.h:1-12): `typedef Phong Lambert`, declared in Material.h here.
#ifndef MIROHOST_LAMBERT_H
#define MIROHOST_LAMBERT_H
#include <cmath>
#include "Object.h"
#include "Material.h"
#endif
