// placeholder: register-tiled Q2 kernels are added here
#pragma once
