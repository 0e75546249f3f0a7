/* TEST INFRASTRUCTURE (oracle/_ref build only) -- not part of the product.
 *
 * Token numbering header that the reference's own sources expect from `yacc -d`
 * (reference src/stcsp.y:1-7,39-41 declares the %token lines).  yacc/bison do not
 * exist in this image, so the header is written by hand: the reference only needs
 * the names to be distinct integers above 255 (single characters such as '<' are
 * used as tokens directly, reference src/node.cpp:93).
 */
#ifndef STCSP_REFSHIM_Y_TAB_H
#define STCSP_REFSHIM_Y_TAB_H

enum {
    STATEMENT = 257, RANGE, LIST,
    VAR, OBJ, ARR,
    LE_CON, GE_CON, EQ_CON, NE_CON, IMPLY_CON, UNTIL_CON,
    LT_OP, GT_OP, LE_OP, GE_OP, EQ_OP, NE_OP,
    AND_OP, OR_OP, NOT_OP,
    AT, FIRST, NEXT, FBY, IF, THEN, ELSE,
    ABS,
    IDENTIFIER, ARR_IDENTIFIER, CONSTANT
};

#endif
