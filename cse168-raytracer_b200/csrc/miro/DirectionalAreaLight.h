// DirectionalAreaLight.h -- forwarding header (reference DirectionalAreaLight.h): the class lives in PointLight.h here.
#ifndef MIROHOST_DIRECTIONALAREALIGHT_H
#define MIROHOST_DIRECTIONALAREALIGHT_H
#include "SquareLight.h"
#endif
