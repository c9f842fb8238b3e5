// SquareLight.h -- forwarding header (reference SquareLight.h): class SquareLight lives in PointLight.h of the host API layer.
#ifndef MIROHOST_SQUARELIGHT_H
#define MIROHOST_SQUARELIGHT_H
#include "PointLight.h"
#endif
