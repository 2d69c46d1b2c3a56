// drop-in for the reference header of the same name: see cloth_compat.h
#include "cloth_compat.h"
