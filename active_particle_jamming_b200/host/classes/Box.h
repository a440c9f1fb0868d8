// One cell of the cell list -- host mirror with the reference's field names (reference
// code/classes/Box.h:4-14). On the device the grid is the counting-sort cell table
// (cell_start / cell_count, csrc/apj_rebuild.cu); Engine::topology fills the geometry and the
// 3x3 neighbour ids here, Engine::pull_cell_lists() materialises CellList when a caller asks.
#ifndef APJ_HOST_BOX_H
#define APJ_HOST_BOX_H

#include <vector>

struct Box
{
    Box();

    int serial_index;
    vector<int> vector_index;
    vector<double> min, max;
    vector<double> center;
    vector<int> neighbors;
    vector<int> CellList;
};

inline Box::Box()
    : serial_index(-1), vector_index(NDIM, -1), min(NDIM, 0), max(NDIM, 0), center(NDIM, 0), neighbors(9, 0)
{
    CellList.reserve(50);
}

#endif
