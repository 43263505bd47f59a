// Stand-in for <GL/gl.h>: cuda_gl_interop.h and cuda_code.cu only need these names.
#pragma once
typedef unsigned int GLuint;
typedef unsigned int GLenum;
#ifndef GL_TEXTURE_3D
#define GL_TEXTURE_3D 0x806F
#endif
