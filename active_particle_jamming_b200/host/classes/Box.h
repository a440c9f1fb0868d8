// One cell of the cell list -- host mirror with the reference's field names (reference
// code/classes/Box.h:4-14). On the device the grid is the counting-sort cell table
// (cell_start / cell_count, csrc/apj_rebuild.cu); Engine::topology fills the geometry and the
// 3x3 neighbour ids here, Engine::pull_cell_lists() materialises CellList when a caller asks.
#ifndef APJ_HOST_BOX_H
#define APJ_HOST_BOX_H

#include <vector>

struct Box
{
    int serial_index = -1;                                   // p = i + j*b (reference numbering)
    vector<int> vector_index = vector<int>(NDIM, -1);        // (i, j)
    vector<double> min = vector<double>(NDIM, 0.0);          // lower corner
    vector<double> max = vector<double>(NDIM, 0.0);          // upper corner
    vector<double> center = vector<double>(NDIM, 0.0);
    vector<int> neighbors = vector<int>(9, 0);               // the 3x3 periodic neighbourhood, 2D
    vector<int> CellList;                                    // particles in the box, ascending index (on request)

    Box() { CellList.reserve(50); }
};

#endif
