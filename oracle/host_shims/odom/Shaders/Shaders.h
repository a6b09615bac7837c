/* TEST INFRASTRUCTURE ONLY: RGBDOdometry.h includes Shaders/Shaders.h (Pangolin GLSL programs) without using anything of it. */
#pragma once
