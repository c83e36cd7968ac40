// TEST INFRASTRUCTURE — no-op stand-in for GL/glew.h so the reference's SPHSystem.cpp and
// Geometry.cpp compile and run without a display. Every GL entry point used by those two files
// is an inline function that does nothing; buffer "names" are handed out from a counter.
#pragma once
#include <cstddef>

typedef unsigned int GLuint;
typedef int GLint;
typedef int GLsizei;
typedef unsigned int GLenum;
typedef unsigned char GLboolean;
typedef float GLfloat;
typedef ptrdiff_t GLsizeiptr;
typedef char GLchar;

#define GL_FALSE 0
#define GL_TRUE 1
#define GL_FLOAT 0x1406
#define GL_UNSIGNED_INT 0x1405
#define GL_TRIANGLES 0x0004
#define GL_ARRAY_BUFFER 0x8892
#define GL_ELEMENT_ARRAY_BUFFER 0x8893
#define GL_STATIC_DRAW 0x88E4
#define GL_DYNAMIC_DRAW 0x88E8
#define GL_WRITE_ONLY 0x88B9

namespace headless_gl {
inline GLuint next_name() { static GLuint n = 0; return ++n; }
}

inline void glGenBuffers(GLsizei n, GLuint *b) { for (GLsizei i = 0; i < n; ++i) b[i] = headless_gl::next_name(); }
inline void glGenVertexArrays(GLsizei n, GLuint *b) { for (GLsizei i = 0; i < n; ++i) b[i] = headless_gl::next_name(); }
inline void glDeleteBuffers(GLsizei, const GLuint *) {}
inline void glDeleteVertexArrays(GLsizei, const GLuint *) {}
inline void glBindBuffer(GLenum, GLuint) {}
inline void glBindVertexArray(GLuint) {}
inline void glBufferData(GLenum, GLsizeiptr, const void *, GLenum) {}
inline void glEnableVertexAttribArray(GLuint) {}
inline void glVertexAttribPointer(GLuint, GLint, GLenum, GLboolean, GLsizei, const void *) {}
inline void glVertexAttribDivisor(GLuint, GLuint) {}
inline void *glMapBuffer(GLenum, GLenum) { return nullptr; }
inline GLboolean glUnmapBuffer(GLenum) { return GL_TRUE; }
inline void glUseProgram(GLuint) {}
inline GLint glGetUniformLocation(GLuint, const GLchar *) { return 0; }
inline void glUniformMatrix4fv(GLint, GLsizei, GLboolean, const GLfloat *) {}
inline void glDrawElements(GLenum, GLsizei, GLenum, const void *) {}
inline void glDrawElementsInstanced(GLenum, GLsizei, GLenum, const void *, GLsizei) {}
