// TEST INFRASTRUCTURE — empty stand-in for GL/glut.h (the simulation step uses nothing from it).
#pragma once
