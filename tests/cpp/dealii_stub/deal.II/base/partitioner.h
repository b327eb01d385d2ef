#include "../stub.h"
